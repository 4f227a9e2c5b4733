"""The read-ahead reader keeps its packed records in reference-counted slabs (seqio.c): records are released one by one, from
other threads, in any order, and a stream may be closed with blocks still queued.  This builds the reader with
AddressSanitizer + UndefinedBehaviorSanitizer + LeakSanitizer and drives exactly that (tests/aux/reader_asan_main.c)."""
import os
import shutil
import subprocess
import numpy as np
import pytest
import _libs as L


def test_slab_records_under_asan(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    exe = str(tmp_path / "reader_asan")
    host = os.path.join(L.ROOT, "airlift_b200", "host")
    cmd = ["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I" + os.path.join(L.ROOT, "include"), "-I" + host,
           os.path.join(L.ROOT, "tests", "aux", "reader_asan_main.c"), os.path.join(host, "seqio.c"), os.path.join(host, "misc.c"), "-o", exe, "-lz", "-lpthread", "-lm"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "asan" in (b.stderr or "").lower():
        pytest.skip("no sanitizer runtime in this toolchain")
    assert b.returncode == 0, b.stderr[-800:]
    rng = np.random.default_rng(12)
    files = []
    for m in (1, 2):
        fn = str(tmp_path / f"r{m}.fq")
        with open(fn, "w") as f:
            for i in range(40000):
                l = int(rng.integers(30, 260))
                s = L.rand_seq(rng, l, 0.01).decode()
                f.write(f"@p{i}/{m} c{i}\n{s}\n+\n{'I' * l}\n")
        files.append(fn)
    r = subprocess.run([exe] + files, capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"), timeout=600)
    assert r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-1500:]
    assert "rep 0: 80000 records" in r.stderr
