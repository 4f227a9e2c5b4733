"""Stage-level parity of the post-chaining device code (airlift_b200/csrc/mmg_{hits,aln,post}.h) on the CPU: the same
functions the kernels of mmg_post.cu call are compiled by g++ (tests/emu/emu.cpp) and run one fragment at a time -- hits,
primary/secondary tree, per-mate split, region planning, CIGAR stitching and clean-up, z-drop cuts, filters, MAPQ, pairing --
and every field of every hit is compared with what the reference's own mm_map_frag (oracle/_ref/libmm2ref.so, map.c:272-424 in
its Oracle-B form) returns for the same fragment.  Chains come from the pinned oracle port (sketch, seeding, mm_chain_dp)."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import _libs as L
import bench
from test_gpu_kernels import _oracle_frag
from test_oracle_vs_ref import _frags, _mk_ref

EMU_SO = os.path.join(L.ROOT, "build", "libmmg_emu.so")

HitRec = np.dtype([("id", "<i4"), ("cnt", "<i4"), ("rid", "<i4"), ("score", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("rs", "<i4"), ("re", "<i4"),
                   ("parent", "<i4"), ("subsc", "<i4"), ("as", "<i4"), ("mlen", "<i4"), ("blen", "<i4"), ("n_sub", "<i4"), ("score0", "<i4"),
                   ("bits", "<u4"), ("hash", "<u4"), ("div", "<f4"), ("p", "<u8")])
assert HitRec.itemsize == 80


class HitOpt(C.Structure):  # airlift_b200/csrc/mmg_hits.h
    _fields_ = [("flag", C.c_int64), ("mask_level", C.c_float), ("pri_ratio", C.c_float), ("max_clip_ratio", C.c_float)] + \
               [(n, C.c_int32) for n in ("best_n", "a", "b", "q", "e", "q2", "e2", "sc_ambi", "zdrop", "zdrop_inv", "end_bonus", "min_dp_max", "min_cnt",
                                         "min_chain_score", "bw", "pe_ori", "pe_bonus", "max_gap", "max_gap_ref", "max_frag_len", "k", "max_qlen")] + \
               [("max_sw_mat", C.c_int64)]


def hit_opt(o):
    h = HitOpt()
    for n, _ in HitOpt._fields_:
        if n != "k":
            setattr(h, n, getattr(o, n))
    return h


@pytest.fixture(scope="module")
def env():
    if not L.have_ref():
        pytest.skip("oracle/_ref/libmm2ref.so not built")
    subprocess.check_call(["make", "-C", L.ROOT, "emu"], stdout=subprocess.DEVNULL)
    E = C.CDLL(EMU_SO)
    E.emu_post_frag.restype = C.c_int
    E.emu_post_frag.argtypes = [C.POINTER(HitOpt), C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    R = L.ref()
    R.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(bench.IdxOpt), C.POINTER(bench.MapOptFull)]
    R.mm_mapopt_update.argtypes = [C.POINTER(bench.MapOptFull), C.c_void_p]
    R.mm_tbuf_init.restype = C.c_void_p
    R.mm_tbuf_destroy.argtypes = [C.c_void_p]
    R.mm_map_frag.restype = None
    R.mm_map_frag.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_void_p,
                              C.POINTER(bench.MapOptFull), C.c_char_p]
    return E, R


def frag_hash(name, qlen_sum, seed):
    """the per-fragment salt of map.c:291-293"""
    def wang(k):
        k = (k + ~(k << 15)) & 0xffffffff; k ^= k >> 10; k = (k + (k << 3)) & 0xffffffff
        k ^= k >> 6; k = (k + ~(k << 11)) & 0xffffffff; k ^= k >> 16
        return k
    h = 0
    if name:
        h = name[0]
        for ch in name[1:]:
            h = ((h << 5) - h + ch) & 0xffffffff
    h ^= (wang(qlen_sum) + wang(seed)) & 0xffffffff
    return wang(h)


def ref_map_frag(R, mi, opt, segs, name):
    n = len(segs)
    qlens = (C.c_int * n)(*[len(s) for s in segs])
    seqs = (C.c_char_p * n)(*segs)
    n_regs = (C.c_int * n)()
    regs = (C.c_void_p * n)()
    tb = R.mm_tbuf_init()
    R.mm_map_frag(mi, n, qlens, seqs, n_regs, regs, tb, C.byref(opt), name)
    R.mm_tbuf_destroy(tb)
    out = []
    for s in range(n):
        hits = []
        if n_regs[s]:
            arr = np.frombuffer(C.string_at(regs[s], n_regs[s] * 80), dtype=HitRec).copy()
            for h in arr:
                ext = None
                if h["p"]:
                    hdr = np.frombuffer(C.string_at(int(h["p"]), 24), dtype="<u4")
                    cig = np.frombuffer(C.string_at(int(h["p"]) + 24, int(hdr[5]) * 4), dtype="<u4").copy()
                    ext = (hdr[1:5].astype(np.int64).tolist(), cig.tolist())  # dp_score, dp_max, dp_max2, n_ambi|strand; cigar (capacity is an allocation size)
                    L.libc.free(int(h["p"]))
                hits.append((h, ext))
            L.libc.free(regs[s])
        out.append(hits)
    return out


def emu_map_frag(E, hopt, k, hsh, segs, refcodes, ref_len, u, a, rep):
    n = len(segs)
    qlens = np.array([len(s) for s in segs], dtype=np.int32)
    codes = [np.ascontiguousarray(L.nt4(s)) for s in segs]
    rp = (C.c_void_p * n)(*[c.ctypes.data for c in codes])
    flip = np.zeros(n, dtype=np.uint8)
    fp = (C.c_void_p * len(refcodes))(*[c.ctypes.data for c in refcodes])
    cap = 256
    n_regs = np.zeros(n, dtype=np.int32)
    regs = np.zeros(n * cap, dtype=HitRec)
    xw = np.zeros(1 << 16, dtype=np.uint32)
    used = C.c_int64(0)
    u = np.ascontiguousarray(u, dtype=np.uint64)
    a = np.ascontiguousarray(a)
    rc = E.emu_post_frag(C.byref(hopt), k, hsh, n, qlens.ctypes.data, rp, flip.ctypes.data, len(refcodes), fp, ref_len.ctypes.data, len(u), u.ctypes.data,
                         a.ctypes.data, rep, n_regs.ctypes.data, regs.ctypes.data, cap, xw.ctypes.data, len(xw), C.byref(used))
    assert rc == 0, rc
    out = []
    for s in range(n):
        hits = []
        for h in regs[s * cap:s * cap + n_regs[s]]:
            ext = None
            if h["p"]:
                o = int(h["p"]) - 1
                hdr = xw[o:o + 6]
                ext = (hdr[1:5].astype(np.int64).tolist(), xw[o + 6:o + 6 + int(hdr[5])].tolist())
            hits.append((h, ext))
        out.append(hits)
    return out


FIELDS = [n for n in HitRec.names if n not in ("p",)]


def same(want, got, ctx):
    assert len(want) == len(got), ctx
    for s, (wl, gl) in enumerate(zip(want, got)):
        assert len(wl) == len(gl), (ctx, s, len(wl), len(gl))
        for i, ((wh, we), (gh, ge)) in enumerate(zip(wl, gl)):
            for f in FIELDS:
                assert wh[f] == gh[f] or (f == "div" and np.isnan(wh[f]) == np.isnan(gh[f])), (ctx, s, i, f, wh[f], gh[f], wh, gh)
            assert (we is None) == (ge is None), (ctx, s, i)
            if we is not None:
                assert we == ge, (ctx, s, i, we, ge)


def make_world(rng, variant):
    refseqs = _mk_ref(rng, n_ctg=2, ln=30000)
    if variant == "dup":  # whole-read duplications: many chains per fragment, secondaries, MAPQ 0
        s = bytearray(refseqs[0])
        for j in range(6):
            o = int(rng.integers(3000, 25000))
            s[o:o + 700] = L.mutate(rng, bytes(refseqs[1][5000:5700]), 0.01 * j, 0, 0)[:700].ljust(700, b"C")
        refseqs[0] = bytes(s)
    return refseqs


@pytest.mark.parametrize("variant,cigar,warp", [("plain", True, 0), ("plain", False, 0), ("dup", True, 0), ("noisy", True, 0), ("zdrop", True, 0), ("single", True, 0),
                                                ("plain", True, 2), ("dup", True, 2), ("single", True, 2)])
def test_post_stages_match_mm_map_frag(env, variant, cigar, warp, monkeypatch):
    """warp > 0: fragments with at least that many chains take the warp-cooperative tree construction (hit_set_parent_warp, the form
    the device uses for fragments from high-copy repeats), its 32 lanes emulated one after the other"""
    E, R = env
    if warp:
        monkeypatch.setenv("EMU_WARP", str(warp))
    else:
        monkeypatch.delenv("EMU_WARP", raising=False)
    rng = np.random.default_rng({"plain": 5, "dup": 6, "noisy": 7, "single": 8, "zdrop": 9}[variant] + (0 if cigar else 100))
    refseqs = make_world(rng, variant)
    w, k = 11, 21
    mi = R.mm_idx_str(w, k, 0, 14, len(refseqs), L.c_str_array(refseqs), None)
    oi = L.oracle().orc_idx_build(w, k, 0, len(refseqs), L.c_str_array(refseqs))
    ipt, opt = bench.IdxOpt(), bench.MapOptFull()
    R.mm_set_opt(None, C.byref(ipt), C.byref(opt))
    R.mm_set_opt(b"sr", C.byref(ipt), C.byref(opt))
    if cigar:
        opt.flag |= 0x004 | 0x008
    R.mm_mapopt_update(C.byref(opt), mi)
    hopt = hit_opt(opt)
    refcodes = [np.ascontiguousarray(L.nt4(s)) for s in refseqs]
    ref_len = np.array([len(s) for s in refseqs], dtype=np.uint32)
    try:
        n_frag = 220
        frags = _frags(rng, refseqs, n_frag, variant != "single")
        if variant == "single":
            frags = [[f[0][:int(rng.integers(60, 400))]] for f in frags]
        if variant == "dup":  # reads from the duplicated block
            for i in range(0, n_frag, 2):
                o = int(rng.integers(5000, 5300))
                frag = refseqs[1][o:o + 380]
                frags[i] = [L.mutate(rng, frag[:150], 0.01, 0.002, 0.002), L.mutate(rng, frag[-150:], 0.01, 0.002, 0.002)]
        if variant == "noisy":  # heavy errors, N runs, long indels: z-drops, clipped extensions, failed filters
            for i in range(n_frag):
                m = []
                for s in frags[i]:
                    s = bytearray(L.mutate(rng, s, 0.06, 0.01, 0.01))
                    if i % 3 == 0:
                        o = int(rng.integers(10, max(11, len(s) - 30))); s[o:o + int(rng.integers(1, 25))] = b"N" * 8
                    if i % 4 == 1 and len(s) > 100:
                        o = int(rng.integers(40, len(s) - 40)); del s[o:o + int(rng.integers(5, 30))]
                    if i % 4 == 2 and len(s) > 100:
                        o = int(rng.integers(40, len(s) - 40)); s[o:o] = L.rand_seq(rng, int(rng.integers(5, 40)))
                    if i % 7 == 3 and len(s) > 120:  # a junction: the tail comes from elsewhere
                        o2 = int(rng.integers(0, 20000)); s[90:] = refseqs[1][o2:o2 + 70]
                    m.append(bytes(s))
                frags[i] = m
        if variant == "zdrop":  # clean reads with one block of mismatches: the ungapped stretch z-drops and the hit is cut (align.c:741-754)
            tr = bytes.maketrans(b"ACGT", b"CATG")
            for i in range(n_frag):
                m = []
                for s in frags[i]:
                    s = bytearray(s)
                    if len(s) > 140 and i % 5 != 4:
                        lo, ln = int(rng.integers(40, 70)), int(rng.integers(12, 45))
                        s[lo:lo + ln] = bytes(s[lo:lo + ln]).translate(tr)
                    m.append(bytes(s))
                frags[i] = m
        n_hits = n_cig = n_split = 0
        for f, segs in enumerate(frags):
            name = b"frag%d" % f
            want = ref_map_frag(R, mi, opt, segs, name)
            u, a, rep, mp, rech = _oracle_frag(oi, opt, segs, w, k, len(segs))
            qsum = sum(len(s) for s in segs)
            got = emu_map_frag(E, hopt, k, frag_hash(name, qsum, opt.seed), segs, refcodes, ref_len, u, a, rep)
            same(want, got, (variant, f))
            n_hits += sum(len(x) for x in want)
            n_cig += sum(1 for x in want for h in x if h[1] is not None)
            n_split += sum(1 for x in want for h in x if (int(h[0]["bits"]) >> 8) & 3)
        assert n_hits > n_frag // 2
        assert (n_cig > 0) == cigar
        if variant == "zdrop":
            assert n_split > 0  # z-drop cuts (mm_split_reg) were exercised
    finally:
        R.mm_idx_destroy(mi)
        L.oracle().orc_idx_destroy(oi)
