"""Multi-GPU plumbing (one process per GPU, torch.distributed): contiguous shard ranges, the one-time index
broadcast, and the ordered merge.  Nothing here sits on the per-batch path: fragments are independent given the
read-only index (reference map.c:458-498), so ranks never exchange data while mapping."""
import ctypes as C
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [begin, end) of rank `rank`; sizes differ by at most one and concatenate to range(n_items)."""
    base, extra = divmod(n_items, world)
    b = rank * base + min(rank, extra)
    return b, b + base + (1 if rank < extra else 0)


class ImageField:
    names = ("S", "seq_off", "seq_len", "slots", "pos")


class IdxImage(C.Structure):  # mmg_idx_image_t (include/mmg.h)
    _fields_ = [("w", C.c_int32), ("k", C.c_int32), ("is_hpc", C.c_int32), ("n_seq", C.c_int32), ("total_len", C.c_uint64),
                ("n_slots", C.c_uint64), ("n_pos", C.c_uint64), ("d_S", C.c_void_p), ("d_seq_off", C.c_void_p),
                ("d_seq_len", C.c_void_p), ("d_slots", C.c_void_p), ("d_pos", C.c_void_p), ("bytes_S", C.c_size_t),
                ("bytes_seq_off", C.c_size_t), ("bytes_seq_len", C.c_size_t), ("bytes_slots", C.c_size_t),
                ("bytes_pos", C.c_size_t), ("n_keys", C.c_int64)]

    def shape_tuple(self):
        return (self.w, self.k, self.is_hpc, self.n_seq, self.total_len, self.n_slots, self.n_pos, self.bytes_S, self.bytes_seq_off,
                self.bytes_seq_len, self.bytes_slots, self.bytes_pos, self.n_keys)

    @classmethod
    def from_shape(cls, t):
        im = cls()
        (im.w, im.k, im.is_hpc, im.n_seq, im.total_len, im.n_slots, im.n_pos, im.bytes_S, im.bytes_seq_off, im.bytes_seq_len,
         im.bytes_slots, im.bytes_pos, im.n_keys) = t
        return im


def broadcast_object(obj, src=0):
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def broadcast_buffers(lib, img: IdxImage, device, src=0, chunk=1 << 30):
    """Broadcast the five device buffers of an index image from `src` to every rank (NCCL over NVLink/NVSwitch).
    The buffers are staged through torch tensors in <=1 GiB pieces (D2D copies on each side)."""
    rank = dist.get_rank()
    total = 0
    for name in ImageField.names:
        ptr = getattr(img, "d_" + name)
        nbytes = getattr(img, "bytes_" + name)
        off = 0
        while off < nbytes:
            n = min(chunk, nbytes - off)
            t = torch.empty(n, dtype=torch.uint8, device=device)
            if rank == src:
                assert lib.mmg_memcpy_d2d(C.c_void_p(t.data_ptr()), C.c_void_p(ptr + off), C.c_size_t(n)) == 0
            dist.broadcast(t, src=src)
            if rank != src:
                assert lib.mmg_memcpy_d2d(C.c_void_p(ptr + off), C.c_void_p(t.data_ptr()), C.c_size_t(n)) == 0
            off += n
            total += n
    return total


def gather_in_order(local_items, dst=0):
    """Ordered merge: rank r holds the results of its contiguous shard; rank `dst` gets them concatenated in input order."""
    world = dist.get_world_size()
    box = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local_items, box, dst=dst)
    if dist.get_rank() != dst:
        return None
    out = []
    for part in box:
        out.extend(part)
    return out
