"""The per-fragment library API (minimap.h:317-334: mm_map_frag / mm_map, regs owned by the caller) on the GPU:
every read pair mapped through mm_map_frag must give the hits the batch path gives (which the end-to-end tests
compare with the reference fork line by line)."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import _libs as L

pytestmark = pytest.mark.gpu
SYN = os.path.join(L.ROOT, "build", "mmsynth")


class Extra(C.Structure):
    _fields_ = [("capacity", C.c_uint32), ("dp_score", C.c_int32), ("dp_max", C.c_int32), ("dp_max2", C.c_int32),
                ("n_ambi_ts", C.c_uint32), ("n_cigar", C.c_uint32)]


class Reg1(C.Structure):  # mm_reg1_t, minimap.h:83-98
    _fields_ = [(n, C.c_int32) for n in ("id", "cnt", "rid", "score", "qs", "qe", "rs", "re", "parent", "subsc", "as_", "mlen", "blen", "n_sub", "score0")] + \
               [("bits", C.c_uint32), ("hash", C.c_uint32), ("div", C.c_float), ("p", C.POINTER(Extra))]


class StepHead(C.Structure):
    _fields_ = [("n_seq", C.c_int), ("n_frag", C.c_int), ("seq", C.c_void_p), ("n_reg", C.POINTER(C.c_int)), ("seg_off", C.POINTER(C.c_int)),
                ("n_seg", C.POINTER(C.c_int)), ("rep_len", C.POINTER(C.c_int)), ("frag_gap", C.POINTER(C.c_int)), ("reg", C.POINTER(C.POINTER(Reg1)))]


def _hit(r, flip_qlen=None):
    qs, qe, rev = r.qs, r.qe, r.bits >> 10 & 1
    if flip_qlen is not None:  # mate mapped as its reverse complement: back to the read's own orientation (map.c:486-497)
        qs, qe, rev = flip_qlen - r.qe, flip_qlen - r.qs, 1 - rev
    cig = []
    if r.p:
        n = r.p.contents.n_cigar
        arr = C.cast(C.addressof(r.p.contents) + C.sizeof(Extra), C.POINTER(C.c_uint32))
        cig = [arr[i] for i in range(n)]
    return (r.rid, r.rs, r.re, qs, qe, rev, r.bits & 0xff, r.score, r.p.contents.dp_max if r.p else -1, tuple(cig))


def test_mm_map_frag_equals_batch_path(tmp_path):
    import bench
    if not os.path.exists(SYN):
        pytest.skip("build/mmsynth missing")
    d = tmp_path
    subprocess.check_call([SYN, "ref", str(d / "ref.fa"), "2000000", "2", "42"])
    subprocess.check_call([SYN, "sr", str(d / "ref.fa"), str(d / "r1.fq"), str(d / "r2.fq"), "60", "44"])
    lib = bench.load_lib()
    lib.mm_map_frag.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.POINTER(Reg1)),
                                C.c_void_p, C.POINTER(bench.MapOptFull), C.c_char_p]
    lib.mm_tbuf_init.restype = C.c_void_p
    lib.mm_tbuf_destroy.argtypes = [C.c_void_p]
    ipt, opt = bench.IdxOpt(), bench.MapOptFull()
    lib.mm_set_opt(None, C.byref(ipt), C.byref(opt))
    lib.mm_set_opt(b"sr", C.byref(ipt), C.byref(opt))
    opt.flag |= 0x004 | 0x008
    rd = lib.mm_idx_reader_open(str(d / "ref.fa").encode(), C.byref(ipt), None)
    mi = lib.mm_idx_reader_read(rd, 3)
    lib.mm_idx_reader_close(rd)
    lib.mm_mapopt_update(C.byref(opt), mi)
    # batch path
    fns = (C.c_char_p * 2)(str(d / "r1.fq").encode(), str(d / "r2.fq").encode())
    reader = lib.mm_b200_open_reads(2, fns)
    b = lib.mm_b200_read_batch(reader, C.byref(opt), 10**8)
    assert lib.mm_b200_map_batch(mi, C.byref(opt), 4, b, 0) == 0
    h = C.cast(b, C.POINTER(StepHead)).contents
    want = {}
    for r in range(h.n_seq):
        want[r] = [_hit(h.reg[r][i]) for i in range(h.n_reg[r])]
    # per-fragment API: sequences in mapping orientation (worker_for flips the second mate of an FR pair, map.c:467-469)
    def reads(fn):
        ls = open(fn).read().split("\n")
        return [(ls[i][1:].split()[0], ls[i + 1]) for i in range(0, len(ls) - 3, 4)]
    r1, r2 = reads(d / "r1.fq"), reads(d / "r2.fq")
    tb = lib.mm_tbuf_init()
    n_hits = 0
    for k, ((n1, s1), (n2, s2)) in enumerate(zip(r1, r2)):
        s2rc = L.revcomp(s2.encode()).decode()
        seqs = (C.c_char_p * 2)(s1.encode(), s2rc.encode())
        qlens = (C.c_int * 2)(len(s1), len(s2))
        n_regs = (C.c_int * 2)()
        regs = (C.POINTER(Reg1) * 2)()
        lib.mm_map_frag(mi, 2, qlens, seqs, n_regs, regs, tb, C.byref(opt), n1.encode())
        got0 = [_hit(regs[0][i]) for i in range(n_regs[0])]
        got1 = [_hit(regs[1][i], flip_qlen=len(s2)) for i in range(n_regs[1])]
        assert got0 == want[2 * k], f"pair {k} mate 1"
        assert got1 == want[2 * k + 1], f"pair {k} mate 2"
        n_hits += len(got0) + len(got1)
    assert n_hits > 60
    lib.mm_tbuf_destroy(tb)
    lib.mm_b200_free_batch(b)
    lib.mm_b200_close_reads(reader)
    lib.mm_idx_destroy(mi)
