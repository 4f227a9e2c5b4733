"""Generates the golden fixtures of tests/golden/ from the REFERENCE itself (not from the oracle):

  kernels.npz   per-function input/output vectors from oracle/_ref/libmm2ref.so (the unmodified reference objects):
                mm_sketch, collect_seed_hits(_heap), mm_chain_dp, ksw_extd2_sse
  cli/          small inputs (made by tools/mmsynth.c and this script) and the output of oracle/_ref/minimap2_B on them

Run here (the container that has /root/reference and the built oracle/_ref):  python tests/golden/make_golden.py
The fixtures travel with the repo, so the oracle and the CUDA path stay pinned where oracle/_ref does not exist."""
import gzip
import json
import os
import subprocess
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs as L                                     # noqa: E402
from test_oracle_vs_ref import _mk_ref, _frags, _ksw_cases, SR_CHAIN, ONT_CHAIN   # noqa: E402

SYN = os.path.join(L.ROOT, "build", "mmsynth")


def u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


def pairs(a):
    return np.ascontiguousarray(a).view(np.uint64).reshape(-1, 2).copy()


def kernels():
    rng = np.random.default_rng(20261017)
    G = {}
    n = 0
    for w, k, hpc in [(11, 21, 0), (10, 15, 0), (5, 19, 1), (3, 4, 0), (1, 15, 0)]:
        for it in range(6):
            ln = int(rng.integers(1, 700))
            s = L.rand_seq(rng, ln, [0, 0.01, 0.2][it % 3])
            if it == 4:
                s = (b"AT" * ln)[:ln]
            if it == 5:
                s = (s[:7] * ln)[:ln]
            G[f"sk{n}_seq"], G[f"sk{n}_par"], G[f"sk{n}_out"] = u8(s), np.array([w, k, 5, hpc]), pairs(L.ref_sketch(s, w, k, 5, hpc))
            n += 1
    G["n_sk"] = np.array(n)

    seqs = _mk_ref(rng, 3, 20000)
    for i, s in enumerate(seqs):
        G[f"ref{i}"] = u8(s)
    G["n_ref"] = np.array(len(seqs))
    import ctypes as C
    for ii, (w, k) in enumerate([(11, 21), (10, 15)]):   # mm_idx_get (index.c:81) and mm_idx_cal_max_occ (index.c:164)
        mi = L.ref().mm_idx_str(w, k, 0, 14, len(seqs), L.c_str_array(seqs), None)
        keys = set()
        for s in seqs:
            keys.update((L.ref_sketch(s[:4000], w, k)["x"] >> np.uint64(8)).tolist())
        keys.update(np.random.default_rng(5 + ii).integers(0, 1 << (2 * k), 300).tolist())   # own stream: the vectors below keep theirs
        keys = np.array(sorted(keys), dtype=np.uint64)
        cnt, pos, n1 = [], [], C.c_int(0)
        for key in keys.tolist():
            p = L.ref().mm_idx_get(mi, key, C.byref(n1))
            cnt.append(n1.value)
            pos.extend(p[j] for j in range(n1.value))
        G[f"ix{ii}_par"], G[f"ix{ii}_keys"], G[f"ix{ii}_n"], G[f"ix{ii}_pos"] = np.array([w, k]), keys, np.array(cnt, dtype=np.int32), np.array(pos, dtype=np.uint64)
        G[f"ix{ii}_occ"] = np.array([L.ref().mm_idx_cal_max_occ(mi, C.c_float(f)) for f in (2e-4, 1e-2, 0.2)], dtype=np.int32)
        L.ref().mm_idx_destroy(mi)
    G["n_ix"] = np.array(2)
    n = 0
    for mode in ["sr", "ont"]:
        w, k = (11, 21) if mode == "sr" else (10, 15)
        mi = L.ref().mm_idx_str(w, k, 0, 14, len(seqs), L.c_str_array(seqs), None)
        n_tie = 0
        for segs in _frags(rng, seqs, 40 if mode == "sr" else 5, mode == "sr"):
            mv, qlen = L.frag_minimizers(L.ref_sketch, segs, w, k)
            for max_occ in ([1000, 5] if mode == "sr" else [50]):
                a, rep, mp = L.ref_collect(mi, mode == "sr", 0, max_occ, mv, qlen)
                tie = len(a) > 1 and bool((a["x"][1:] == a["x"][:-1]).any())
                n_tie += tie
                u, b = L.chain_call(L.ref().ref_chain_dp, SR_CHAIN if mode == "sr" else ONT_CHAIN, a)
                for j, sg in enumerate(segs):
                    G[f"fr{n}_seg{j}"] = u8(sg)
                G[f"fr{n}_par"] = np.array([mode == "sr", w, k, max_occ, len(segs), qlen, rep])
                G[f"fr{n}_mv"], G[f"fr{n}_a"], G[f"fr{n}_mp"], G[f"fr{n}_u"], G[f"fr{n}_b"] = pairs(mv), pairs(a), mp, u, pairs(b)
                n += 1
        assert mode != "sr" or n_tie > 0
        L.ref().mm_idx_destroy(mi)
    G["n_fr"] = np.array(n)

    n = 0
    for preset in ["sr", "ont"]:
        if preset == "sr":
            mat, pen, bw, zd, eb = L.simple_mat(2, 8, 1), (12, 2, 24, 1), 151, 100, 10
        else:
            mat, pen, bw, zd, eb = L.simple_mat(2, 4, 1), (4, 2, 24, 1), 751, 400, -1
        for ci, (q, t) in enumerate(_ksw_cases(rng, 36)):
            for fl in [0xC2, 0x40, 0x08, 0x00, 0x02]:
                for w in ([bw, 20] if ci % 6 == 0 else [bw]):
                    r = L.ref_ksw(q, t, mat, *pen, w, zd, eb if fl & 0x40 else -1, fl)
                    G[f"kw{n}_q"], G[f"kw{n}_t"] = q, t
                    G[f"kw{n}_par"] = np.array([preset == "sr", w, fl])
                    G[f"kw{n}_res"] = np.array([r[x] for x in ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end")], dtype=np.int64)
                    G[f"kw{n}_cig"] = np.array(r["cigar"], dtype=np.uint32)
                    n += 1
    G["n_kw"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **G)
    print("kernels.npz:", int(G["n_sk"]), "sketch,", int(G["n_fr"]), "seed/chain,", int(G["n_kw"]), "ksw cases")


CLI_CASES = {   # name -> arguments (relative to cli/); the expected output is cli/<name>.out.gz
    "sr_sam": ["-ax", "sr", "-t", "4", "ref.fa", "r1.fq", "r2.fq"],
    "sr_paf": ["-x", "sr", "-t", "4", "ref.fa", "r1.fq", "r2.fq"],
    "sr_single_cs": ["-ax", "sr", "-t", "4", "--cs", "--MD", "ref.fa", "r1.fq"],
    "sr_lowocc": ["-ax", "sr", "-t", "4", "-f", "10,30", "ref.fa", "r1.fq", "r2.fq"],
    "sr_secondary": ["-ax", "sr", "-t", "4", "--secondary=yes", "-N", "20", "-p", "0.3", "ref.fa", "r1.fq"],
    "sr_edge": ["-ax", "sr", "-t", "2", "ref.fa", "edge.fq"],
    "ont_paf_cs": ["-cx", "map-ont", "-t", "4", "--cs", "ref.fa", "long.fq"],
    "ont_sam": ["-ax", "map-ont", "-t", "4", "ref.fa", "long.fq"],
    "inv_paf": ["-cx", "map-ont", "inv_t.fa", "inv_q.fa"],
}


def cli():
    d = os.path.join(HERE, "cli")
    os.makedirs(d, exist_ok=True)
    subprocess.check_call([SYN, "ref", os.path.join(d, "ref.fa"), "300000", "2", "42"])
    subprocess.check_call([SYN, "sr", os.path.join(d, "ref.fa"), os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), "300", "44", "0.05"])
    subprocess.check_call([SYN, "long", os.path.join(d, "ref.fa"), os.path.join(d, "long.fq"), "8", "45", "6000"])
    ref = "".join(open(os.path.join(d, "ref.fa")).read().split("\n")[1:60])
    s = ref[1000:1150]
    with open(os.path.join(d, "edge.fq"), "w") as f:   # empty-ish, shorter than k, all N, lower case, RNA alphabet
        f.write("@short\nACGTACGT\n+\nIIIIIIII\n@allN\n" + "N" * 150 + "\n+\n" + "I" * 150 + "\n")
        f.write("@lower\n" + s.lower() + "\n+\n" + "I" * 150 + "\n@rna\n" + s.replace("T", "U") + "\n+\n" + "I" * 150 + "\n")
    rng = np.random.default_rng(7)
    t = L.rand_seq(rng, 6000)
    q = L.mutate(rng, t[:2500], 0.02, 0.01, 0.01) + L.revcomp(L.mutate(rng, t[2500:3300], 0.02, 0.01, 0.01)) + L.mutate(rng, t[3300:], 0.02, 0.01, 0.01)
    open(os.path.join(d, "inv_t.fa"), "w").write(">t_inv\n" + t.decode() + "\n")
    open(os.path.join(d, "inv_q.fa"), "w").write(">q_inv\n" + q.decode() + "\n")
    for name, args in CLI_CASES.items():
        p = subprocess.run([L.REF_BIN_B] + args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
        lines = [l for l in p.stdout.decode().split("\n") if not l.startswith("@PG")]
        with gzip.GzipFile(os.path.join(d, name + ".out.gz"), "wb", mtime=0) as f:
            f.write("\n".join(lines).encode())
        print(name, len(lines), "lines")
    with open(os.path.join(d, "cases.json"), "w") as f:
        json.dump(CLI_CASES, f, indent=1)
    for fn in ["ref.fa", "r1.fq", "r2.fq", "long.fq"]:
        with open(os.path.join(d, fn), "rb") as src, gzip.GzipFile(os.path.join(d, fn + ".gz"), "wb", mtime=0) as dst:
            dst.write(src.read())
        os.remove(os.path.join(d, fn))


if __name__ == "__main__":
    assert L.have_ref() and os.path.exists(L.REF_BIN_B), "build oracle/_ref first (make -C oracle all)"
    kernels()
    cli()
