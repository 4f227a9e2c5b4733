"""Golden fixtures (tests/golden/, generated from the reference itself by tests/golden/make_golden.py):

  * CPU: the plain-C oracle reproduces every reference vector -- this pins the oracle where oracle/_ref does not exist;
  * GPU (-m gpu): the CUDA stages through the C-ABI reproduce the same vectors, and the CLI reproduces the
    reference's SAM/PAF for the committed inputs (sr paired/single, low occurrence cut-offs, many secondaries,
    edge-case reads, map-ont with cs, an inversion that needs the tp:A:I path)."""
import ctypes as C
import gzip
import json
import os
import shutil
import subprocess
import numpy as np
import pytest
import _libs as L

GOLD = os.path.join(L.ROOT, "tests", "golden")
SR_CHAIN = (500, 300, 100, 25, 5000, 2, 25, 0, 2)
ONT_CHAIN = (5000, 5000, 500, 25, 5000, 3, 40, 0, 1)
RES_KEYS = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end")


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLD, "kernels.npz"))


def as128(p):
    return np.ascontiguousarray(p).view(L.mm128).reshape(-1)


def _refseqs(G):
    return [G[f"ref{i}"].tobytes() for i in range(int(G["n_ref"]))]


def _ksw_params(sr):
    if sr:
        return L.simple_mat(2, 8, 1), (12, 2, 24, 1), 100, 10
    return L.simple_mat(2, 4, 1), (4, 2, 24, 1), 400, -1


def _check_sketch(G, fn):
    for i in range(int(G["n_sk"])):
        w, k, rid, hpc = (int(x) for x in G[f"sk{i}_par"])
        got = fn(G[f"sk{i}_seq"].tobytes(), w, k, rid, hpc)
        assert got.tobytes() == as128(G[f"sk{i}_out"]).tobytes(), (i, w, k, hpc)


def _check_ksw(G, fn):
    for i in range(int(G["n_kw"])):
        sr, w, fl = (int(x) for x in G[f"kw{i}_par"])
        mat, pen, zd, eb = _ksw_params(sr)
        r = fn(G[f"kw{i}_q"], G[f"kw{i}_t"], mat, *pen, w, zd, eb if fl & 0x40 else -1, fl)
        assert [int(r[x]) for x in RES_KEYS] == [int(x) for x in G[f"kw{i}_res"]], (i, sr, w, fl)
        assert [int(x) for x in r["cigar"]] == [int(x) for x in G[f"kw{i}_cig"]], (i, sr, w, fl)


# ----------------------------------------------------------------- CPU: the oracle against the reference's vectors

def test_oracle_sketch_golden(G):
    _check_sketch(G, L.orc_sketch)


def test_oracle_seeds_and_chain_golden(G):
    seqs = _refseqs(G)
    idx = {}
    try:
        for i in range(int(G["n_fr"])):
            sr, w, k, max_occ, n_segs, qlen, rep = (int(x) for x in G[f"fr{i}_par"])
            if (w, k) not in idx:
                idx[(w, k)] = L.oracle().orc_idx_build(w, k, 0, len(seqs), L.c_str_array(seqs))
            segs = [G[f"fr{i}_seg{j}"].tobytes() for j in range(n_segs)]
            mv, ql = L.frag_minimizers(L.orc_sketch, segs, w, k)
            assert ql == qlen and mv.tobytes() == as128(G[f"fr{i}_mv"]).tobytes()
            a, rep1, mp = L.orc_collect(idx[(w, k)], bool(sr), 0, max_occ, mv, qlen)
            assert rep1 == rep and (mp == G[f"fr{i}_mp"]).all() and a.tobytes() == as128(G[f"fr{i}_a"]).tobytes(), i
            u, b = L.chain_call(L.oracle().orc_chain_dp, SR_CHAIN if sr else ONT_CHAIN, a)
            assert (u == G[f"fr{i}_u"]).all() and b.tobytes() == as128(G[f"fr{i}_b"]).tobytes(), i
    finally:
        for h in idx.values():
            L.oracle().orc_idx_destroy(h)


def test_oracle_index_golden(G):
    seqs = _refseqs(G)
    for ii in range(int(G["n_ix"])):
        w, k = (int(x) for x in G[f"ix{ii}_par"])
        oi = L.oracle().orc_idx_build(w, k, 0, len(seqs), L.c_str_array(seqs))
        try:
            assert [L.oracle().orc_idx_cal_max_occ(oi, f) for f in (2e-4, 1e-2, 0.2)] == G[f"ix{ii}_occ"].tolist()
            n1, at, pos = C.c_int(0), 0, G[f"ix{ii}_pos"]
            for key, n in zip(G[f"ix{ii}_keys"].tolist(), G[f"ix{ii}_n"].tolist()):
                p = L.oracle().orc_idx_get(oi, key, C.byref(n1))
                assert n1.value == n and [p[j] for j in range(n)] == pos[at:at + n].tolist()
                at += n
        finally:
            L.oracle().orc_idx_destroy(oi)


def test_oracle_ksw_golden(G):
    _check_ksw(G, L.orc_ksw)


# ----------------------------------------------------------------- GPU: the CUDA path against the same vectors

@pytest.fixture(scope="module")
def ctx():
    import airlift_b200.api as A
    c = A.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_cuda_sketch_golden(G, ctx):
    _check_sketch(G, ctx.sketch)


@pytest.mark.gpu
def test_cuda_seeds_and_chain_golden(G, ctx):
    import airlift_b200.api as A
    seqs = _refseqs(G)
    idx = {}
    try:
        for i in range(int(G["n_fr"])):
            sr, w, k, max_occ, n_segs, qlen, rep = (int(x) for x in G[f"fr{i}_par"])
            if (w, k) not in idx:
                idx[(w, k)] = A.Index(ctx, seqs, w, k)
            mv, want = as128(G[f"fr{i}_mv"]), as128(G[f"fr{i}_a"])
            a, rep1, mp = ctx.collect_seeds(idx[(w, k)], bool(sr), 0, max_occ, mv, qlen, len(want) + 8)
            assert rep1 == rep and (mp == G[f"fr{i}_mp"]).all() and a.tobytes() == want.tobytes(), i
            u, b = ctx.chain_dp(SR_CHAIN if sr else ONT_CHAIN, want)
            assert (u == G[f"fr{i}_u"]).all() and b.tobytes() == as128(G[f"fr{i}_b"]).tobytes(), i
    finally:
        for h in idx.values():
            h.close()


@pytest.mark.gpu
def test_cuda_index_golden(G, ctx):
    import airlift_b200.api as A
    seqs = _refseqs(G)
    for ii in range(int(G["n_ix"])):
        w, k = (int(x) for x in G[f"ix{ii}_par"])
        idx = A.Index(ctx, seqs, w, k)
        try:
            assert [idx.cal_max_occ(f) for f in (2e-4, 1e-2, 0.2)] == G[f"ix{ii}_occ"].tolist()
            want_n, want_pos = G[f"ix{ii}_n"], G[f"ix{ii}_pos"]
            n, pos = idx.get(G[f"ix{ii}_keys"], int(want_n.max()) + 1)
            assert (np.asarray(n) == want_n).all()
            at = 0
            for i, m in enumerate(want_n.tolist()):
                assert pos[i, :m].tolist() == want_pos[at:at + m].tolist()
                at += m
        finally:
            idx.close()


@pytest.mark.gpu
def test_cuda_ksw_golden(G, ctx):
    _check_ksw(G, ctx.ksw_extd2)


@pytest.fixture(scope="module")
def cli_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_cli")
    src = os.path.join(GOLD, "cli")
    for fn in os.listdir(src):
        if fn.endswith(".out.gz"):
            continue
        if fn.endswith(".gz"):
            with gzip.open(os.path.join(src, fn), "rb") as f, open(d / fn[:-3], "wb") as o:
                shutil.copyfileobj(f, o)
        else:
            shutil.copy(os.path.join(src, fn), d / fn)
    return d


def _cli_cases():
    with open(os.path.join(GOLD, "cli", "cases.json")) as f:
        return json.load(f)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_cli_cases()))
def test_cli_golden(cli_dir, name):
    new = os.path.join(L.ROOT, "build", "minimap2-b200")
    if not os.path.exists(new):
        pytest.skip("build/minimap2-b200 not built")
    p = subprocess.run([new] + _cli_cases()[name], cwd=cli_dir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()[-800:]
    got = [l for l in p.stdout.decode().split("\n") if not l.startswith("@PG")]
    want = gzip.open(os.path.join(GOLD, "cli", name + ".out.gz"), "rb").read().decode().split("\n")
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(want, got)):
        assert a == b, f"line {i}:\nref: {a[:300]}\nnew: {b[:300]}"
