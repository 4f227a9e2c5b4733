"""Row f4 of the scope table: `ksw_extz2_sse` (ksw2_extz2_sse.c:23) is what mm_align_pair calls when both gap-cost pairs are
equal (align.c:328-329).  With q2 == q and e2 == e the second gap state of ksw_extd2_sse evolves exactly like the first
(same initial values, same recurrence), the direction codes 3/4 never win a strict comparison, and the boundary column and
z-drop use the same constants -- so the dual-cost kernel computes the single-cost function.  This test pins that claim
on the reference itself, for every flag combination align.c passes (KSW_EZ_APPROX_DROP is never set by it: align.c:733 is
the only approximate call and sets KSW_EZ_APPROX_MAX alone); the CUDA side is held to it by the `-O x,x -E y,y` CLI cases
of tests/test_gpu_e2e.py."""
import ctypes as C
import numpy as np
import pytest
import _libs as L


def _ref_extz2(q, t, mat, gq, ge, w, zdrop, end_bonus, flag):
    ez = L.RefExtz()
    C.memset(C.byref(ez), 0, C.sizeof(ez))
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    L.ref().ksw_extz2_sse(None, len(q), q.ctypes.data, len(t), t.ctypes.data, 5, mat.ctypes.data, gq, ge, w, zdrop, end_bonus, flag, C.byref(ez))
    r = L._extz_tuple(ez, ez.max_zd >> 31, ez.max_zd & 0x7fffffff)
    L.libc.free(ez.cigar)
    return r


def test_extz2_is_extd2_with_equal_gap_costs():
    if not L.have_ref():
        pytest.skip("needs oracle/_ref/libmm2ref.so")
    from test_oracle_vs_ref import _ksw_cases
    rng = np.random.default_rng(5)
    n = 0
    for mat, (gq, ge), bands, zd in [(L.simple_mat(2, 8, 1), (12, 2), [151, 20, -1], 100), (L.simple_mat(2, 4, 1), (4, 2), [751, 30, -1], 400),
                                     (L.simple_mat(1, 4, 1), (6, 1), [50, 8], 200)]:
        for q, t in _ksw_cases(rng, 60):
            for fl in [0xC2, 0x40, 0x08, 0x00, 0x01, 0x02, 0x48, 0x82, 0x09, 0x4A]:
                for w in bands:
                    eb = 10 if fl & 0x40 else -1
                    assert _ref_extz2(q, t, mat, gq, ge, w, zd, eb, fl) == L.ref_ksw(q, t, mat, gq, ge, gq, ge, w, zd, eb, fl), (len(q), len(t), fl, w)
                    # and the oracle's dual-cost restatement, which is what the CUDA kernels are compared with
                    assert L.orc_ksw(q, t, mat, gq, ge, gq, ge, w, zd, eb, fl) == _ref_extz2(q, t, mat, gq, ge, w, zd, eb, fl) or (fl & 0x01)
                    n += 1
    assert n > 4000
