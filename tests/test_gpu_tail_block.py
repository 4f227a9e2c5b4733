"""The CTA form of the chain tail (k_chain_tail_block) normally takes the few fragments with thousands of anchors.
MMG_TAIL_HEAVY_N=0 sends every fragment of the first pass through it, so the chaining parity tests (oracle, golden vectors,
CLI byte identity incl. the repeat-family fixture) all run on its level-by-level backtracking; MMG_TAIL_WALK=1 selects the
chain-walking form it replaced, which stays as the path for presets with min_cnt < 2.  Both variables are read once per
process, hence the subprocesses."""
import os
import subprocess
import sys
import pytest
import _libs as L

pytestmark = pytest.mark.gpu
SUITE = ["tests/test_gpu_kernels.py", "tests/test_golden.py", "tests/test_gpu_paths.py", "tests/test_gpu_e2e.py"]


@pytest.mark.parametrize("walk", [0, 1])
def test_chaining_suite_through_the_cta_tail(walk):
    env = dict(os.environ, MMG_TAIL_HEAVY_N="0")
    if walk:
        env["MMG_TAIL_WALK"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "chain or golden or paths or e2e or cli"] + SUITE,
                       cwd=L.ROOT, env=env, capture_output=True, text=True, timeout=2400)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1000:])
    assert " passed" in r.stdout
