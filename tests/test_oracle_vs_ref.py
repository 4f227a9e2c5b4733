"""Pins the plain-C oracle (oracle/mm2_oracle.c) to the compiled reference fork
(oracle/_ref/libmm2ref.so built from /root/reference by oracle/Makefile).
CPU-only.  Skipped when the reference build is not available (then the committed
fixtures in tests/golden/ -- generated from this same comparison -- pin it instead)."""
import ctypes as C
import numpy as np
import pytest
import _libs as L

pytestmark = pytest.mark.skipif(not L.have_ref(), reason="oracle/_ref/libmm2ref.so not built")

SR_CHAIN = (500, 300, 100, 25, 5000, 2, 25, 0, 2)
ONT_CHAIN = (5000, 5000, 500, 25, 5000, 3, 40, 0, 1)


@pytest.mark.parametrize("w,k,hpc", [(11, 21, 0), (10, 15, 0), (5, 19, 1), (19, 19, 0), (3, 4, 0), (1, 15, 0), (10, 14, 0)])
def test_sketch(w, k, hpc):
    rng = np.random.default_rng(1000 + w * 31 + k)
    for it in range(60):
        n = int(rng.integers(1, 600))
        s = L.rand_seq(rng, n, n_frac=[0, 0.01, 0.2][it % 3])
        if it % 7 == 0:  # low-complexity: forces identical k-mers / palindromes
            s = (b"AT" * n)[:n] if it % 2 else (s[:7] * n)[:n]
        a, b = L.orc_sketch(s, w, k, 3, hpc), L.ref_sketch(s, w, k, 3, hpc)
        assert a.tobytes() == b.tobytes(), (w, k, hpc, it)


def test_radix_sorts():
    rng = np.random.default_rng(7)
    for n in [0, 1, 2, 63, 64, 65, 200, 5000]:
        for keybits in [3, 8, 20, 64]:
            a = np.zeros(n, dtype=L.mm128)
            hi = (1 << keybits) - 1
            a["x"] = rng.integers(0, hi, n, dtype=np.uint64, endpoint=True)
            a["y"] = np.arange(n)
            b = a.copy()
            L.oracle().orc_radix_sort_128x(a.ctypes.data, a.ctypes.data + 16 * n)
            L.ref().radix_sort_128x(b.ctypes.data, b.ctypes.data + 16 * n)
            assert a.tobytes() == b.tobytes(), (n, keybits)
            u = a["x"].copy(); rng.shuffle(u); v = u.copy()
            L.oracle().orc_radix_sort_64(u.ctypes.data, u.ctypes.data + 8 * n)
            L.ref().radix_sort_64(v.ctypes.data, v.ctypes.data + 8 * n)
            assert (u == v).all() and (np.diff(u.astype(object)) >= 0).all()


def _mk_ref(rng, n_ctg=3, ln=40000):
    seqs = []
    for i in range(n_ctg):
        s = bytearray(L.rand_seq(rng, ln, 0.001))
        rep = s[1000:1400]
        for j in range(30):  # a repeat family so that occurrence cut-offs fire
            o = int(rng.integers(2000, ln - 500))
            s[o:o + 400] = L.mutate(rng, bytes(rep), 0.02, 0, 0)[:400].ljust(400, b"A")
        seqs.append(bytes(s))
    return seqs


@pytest.fixture(scope="module")
def small_ref():
    rng = np.random.default_rng(42)
    return _mk_ref(rng)


@pytest.mark.parametrize("w,k", [(11, 21), (10, 15)])
def test_index_lookup(small_ref, w, k):
    seqs = small_ref
    arr = L.c_str_array(seqs)
    mi = L.ref().mm_idx_str(w, k, 0, 14, len(seqs), arr, None)
    oi = L.oracle().orc_idx_build(w, k, 0, len(seqs), arr)
    try:
        for f in [2e-4, 1e-2, 0.2]:
            assert L.ref().mm_idx_cal_max_occ(mi, f) == L.oracle().orc_idx_cal_max_occ(oi, f)
        rng = np.random.default_rng(5)
        keys = set()
        for s in seqs:
            keys.update((L.orc_sketch(s[:5000], w, k)["x"] >> np.uint64(8)).tolist())
        keys.update(rng.integers(0, 1 << (2 * k), 500).tolist())  # mostly absent
        n1, n2 = C.c_int(0), C.c_int(0)
        for key in keys:
            p1 = L.ref().mm_idx_get(mi, key, C.byref(n1))
            p2 = L.oracle().orc_idx_get(oi, key, C.byref(n2))
            assert n1.value == n2.value
            assert [p1[i] for i in range(n1.value)] == [p2[i] for i in range(n2.value)]
    finally:
        L.ref().mm_idx_destroy(mi)
        L.oracle().orc_idx_destroy(oi)


def _frags(rng, seqs, n, paired):
    out = []
    for i in range(n):
        s = seqs[int(rng.integers(0, len(seqs)))]
        if paired:
            ins = int(np.clip(rng.normal(330, 80), 160, 700))  # many overlapping mates: equal-x anchors
            o = int(rng.integers(0, len(s) - ins))
            frag = s[o:o + ins]
            if rng.random() < 0.5:
                frag = L.revcomp(frag)
            m1 = L.mutate(rng, frag[:150], 0.01, 0.0015, 0.0005)
            m2 = L.mutate(rng, L.revcomp(frag)[:150], 0.01, 0.0015, 0.0005)
            out.append([m1, L.revcomp(m2)])  # mate 2 is reverse-complemented before mapping (map.c:467-469)
        else:
            ln = int(rng.integers(800, 6000))
            o = int(rng.integers(0, len(s) - ln))
            r = s[o:o + ln]
            if rng.random() < 0.5:
                r = L.revcomp(r)
            out.append([L.mutate(rng, r, 0.03, 0.03, 0.03)])
    return out


@pytest.mark.parametrize("mode", ["sr", "ont"])
def test_seeds_and_chain(small_ref, mode):
    seqs = small_ref
    w, k = (11, 21) if mode == "sr" else (10, 15)
    arr = L.c_str_array(seqs)
    mi = L.ref().mm_idx_str(w, k, 0, 14, len(seqs), arr, None)
    oi = L.oracle().orc_idx_build(w, k, 0, len(seqs), arr)
    rng = np.random.default_rng(99)
    n_tie = 0
    try:
        for segs in _frags(rng, seqs, 150 if mode == "sr" else 25, mode == "sr"):
            mv_o, qlen = L.frag_minimizers(L.orc_sketch, segs, w, k)
            mv_r, _ = L.frag_minimizers(L.ref_sketch, segs, w, k)
            assert mv_o.tobytes() == mv_r.tobytes()
            for max_occ in ([1000, 20, 5] if mode == "sr" else [50, 8]):
                a1, rep1, mp1 = L.orc_collect(oi, mode == "sr", 0, max_occ, mv_o, qlen)
                a2, rep2, mp2 = L.ref_collect(mi, mode == "sr", 0, max_occ, mv_o, qlen)
                assert rep1 == rep2 and (mp1 == mp2).all()
                assert a1.tobytes() == a2.tobytes()
                if len(a1) > 1 and (a1["x"][1:] == a1["x"][:-1]).any():
                    n_tie += 1
                params = SR_CHAIN if mode == "sr" else ONT_CHAIN
                u1, b1 = L.chain_call(L.oracle().orc_chain_dp, params, a1)
                u2, b2 = L.chain_call(L.ref().ref_chain_dp, params, a1)
                assert (u1 == u2).all() and b1.tobytes() == b2.tobytes()
        if mode == "sr":
            assert n_tie > 0  # the equal-x tie order case was exercised
    finally:
        L.ref().mm_idx_destroy(mi)
        L.oracle().orc_idx_destroy(oi)


def _ksw_cases(rng, n):
    for i in range(n):
        kind = i % 6
        if kind == 0:    # short-read extension shapes
            ql, tl = int(rng.integers(1, 150)), int(rng.integers(1, 230))
        elif kind == 1:  # band-binding corner (w=151): q<=150, t up to ~225
            ql, tl = int(rng.integers(100, 151)), int(rng.integers(180, 230))
        elif kind == 2:  # ont gap fill
            ql, tl = int(rng.integers(150, 500)), int(rng.integers(150, 500))
        elif kind == 3:  # tiny
            ql, tl = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        elif kind == 4:  # multiples of 16
            ql, tl = 16 * int(rng.integers(1, 12)), 16 * int(rng.integers(1, 12))
        else:
            ql, tl = int(rng.integers(1, 700)), int(rng.integers(1, 700))
        t = L.rand_seq(rng, tl, 0.01)
        if rng.random() < 0.8:
            base = (t + L.rand_seq(rng, ql))[:ql] if ql > tl else t[:ql]
            q = L.mutate(rng, base, 0.05, 0.03, 0.03)
            if rng.random() < 0.3:  # a long gap
                c = int(rng.integers(0, max(1, len(q))))
                q = q[:c] + L.rand_seq(rng, int(rng.integers(5, 60))) + q[c:]
            q = (q + L.rand_seq(rng, ql))[:ql]
        else:
            q = L.rand_seq(rng, ql, 0.02)
        yield L.nt4(q), L.nt4(t)


@pytest.mark.parametrize("preset", ["sr", "ont"])
def test_ksw_extd2(preset):
    rng = np.random.default_rng(2024)
    if preset == "sr":
        mat, pen, bw, zd, eb = L.simple_mat(2, 8, 1), (12, 2, 24, 1), 151, 100, 10
    else:
        mat, pen, bw, zd, eb = L.simple_mat(2, 4, 1), (4, 2, 24, 1), 751, 400, -1
    flags = [0x40 | 0x02 | 0x80, 0x40, 0x08, 0x00, 0x01, 0x02]  # left ext, right ext, approx gap fill, exact gap fill, score only, right-aligned global
    n = 0
    for q, t in _ksw_cases(rng, 240):
        for fl in flags:
            for w in ([bw, 20, -1] if n % 5 == 0 else [bw]):
                a = L.orc_ksw(q, t, mat, *pen, w, zd, eb if fl & 0x40 else -1, fl)
                b = L.ref_ksw(q, t, mat, *pen, w, zd, eb if fl & 0x40 else -1, fl)
                if fl & 0x01:
                    a["cigar"] = b["cigar"] = []
                assert a == b, (len(q), len(t), fl, w)
        n += 1
