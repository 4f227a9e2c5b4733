"""The CTA form of the literal K4 kernel (k_ksw_dpx_block: one CTA per job, lane state in shared memory) only takes jobs whose
state exceeds the warp kernel's slot -- long-read end extensions.  MMG_KSW_FORCE_BLOCK=1 sends every literal job it can
take through it, so the whole single-job parity suite (oracle + golden vectors, all flag / band combinations) runs on it; the
environment variable is read once per process, hence the subprocess.  A second test maps kilobase jobs with a clipped band
in-process and checks that the kernel launched on its own."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np
import pytest
import _libs as L

pytestmark = pytest.mark.gpu


def _launches(ctx, name):
    from airlift_b200 import api
    lib = api.lib()
    names = (C.c_char_p * 256)(); ms = (C.c_double * 256)(); n = (C.c_long * 256)()
    lib.mmg_profile_fetch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_long)]
    k = lib.mmg_profile_fetch(ctx.h, 256, names, ms, n)
    return sum(int(n[i]) for i in range(k) if names[i] and names[i].decode() == name)


def test_single_job_suite_through_the_cta_form():
    env = dict(os.environ, MMG_KSW_FORCE_BLOCK="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "ksw and not cta_form", "tests/test_gpu_kernels.py", "tests/test_golden.py"],
                       cwd=L.ROOT, env=env, capture_output=True, text=True, timeout=2400)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1000:])
    assert " passed" in r.stdout


def test_large_clipped_jobs_take_the_cta_form():
    from airlift_b200 import api
    lib = api.lib()
    ctx = api.Context(0)
    try:
        lib.mmg_profile_enable.argtypes = [C.c_void_p, C.c_int]
        lib.mmg_profile_enable(ctx.h, 1)
        rng = np.random.default_rng(77)
        mat, pen = L.simple_mat(2, 4, 1), (4, 2, 24, 1)
        for ql, tl, w, fl in [(700, 1300, 500, 0xC2), (1500, 2400, 500, 0x40), (900, 900, 200, 0x00), (2500, 3000, 500, 0x4A), (1200, 1100, 100, 0x01)]:
            t = rng.integers(0, 4, tl, dtype=np.uint8)
            q = t[:ql].copy() if ql <= tl else np.concatenate([t, rng.integers(0, 4, ql - tl, dtype=np.uint8)])
            mut = rng.random(ql) < 0.12
            q[mut] = rng.integers(0, 4, int(mut.sum()), dtype=np.uint8)
            q = np.delete(q, rng.integers(0, ql, 12))
            a = L.orc_ksw(q, t, mat, *pen, w, 400, 10 if fl & 0x40 else -1, fl)
            b = ctx.ksw_extd2(q, t, mat, *pen, w, 400, 10 if fl & 0x40 else -1, fl)
            if fl & 0x01:
                a["cigar"] = []
            assert a == b, (ql, tl, w, fl)
        assert _launches(ctx, "k_ksw_dpx_block") >= 3
    finally:
        ctx.close()
