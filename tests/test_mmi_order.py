"""Slot order of the .mmi hash tables (index.c:438-477 writes klib hash tables in slot order).

The host side of mm_idx_dump() replays klib's placement on slot numbers (airlift_b200/host/refidx.c, kh_replay).
Here that replay is checked on the CPU against index files written by the reference itself
(oracle/_ref/minimap2_B -d): for every bucket, inserting the bucket's keys in ascending order (the order
worker_post, index.c:191-243, inserts them) must give the order the file lists them in."""
import ctypes as C
import os
import random
import struct
import subprocess
import numpy as np
import pytest
import _libs as L

LIB = os.path.join(L.ROOT, "airlift_b200", "libmm2b200.so")


def parse_mmi(path):
    """-> list of parts: dict(w,k,b,flag,names,lens,buckets=[(p, keys, vals)], S)"""
    raw = open(path, "rb").read()
    at, parts = 0, []
    while at < len(raw):
        assert raw[at:at + 4] == b"MMI\2"
        w, k, b, n_seq, flag = struct.unpack_from("<5I", raw, at + 4)
        at += 24
        names, lens = [], []
        for _ in range(n_seq):
            l = raw[at]
            names.append(raw[at + 1:at + 1 + l])
            lens.append(struct.unpack_from("<I", raw, at + 1 + l)[0])
            at += 5 + l
        buckets = []
        for _ in range(1 << b):
            n = struct.unpack_from("<i", raw, at)[0]
            p = np.frombuffer(raw, dtype="<u8", count=n, offset=at + 4)
            at += 4 + 8 * n
            size = struct.unpack_from("<I", raw, at)[0]
            kv = np.frombuffer(raw, dtype="<u8", count=2 * size, offset=at + 4).reshape(-1, 2)
            at += 4 + 16 * size
            buckets.append((p, kv[:, 0], kv[:, 1]))
        n_words = (sum(lens) + 7) // 8 if not flag & 2 else 0
        S = np.frombuffer(raw, dtype="<u4", count=n_words, offset=at)
        at += 4 * n_words
        parts.append(dict(w=w, k=k, b=b, flag=flag, names=names, lens=lens, buckets=buckets, S=S))
    return parts


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("libmm2b200.so not built")
    so = C.CDLL(LIB)
    so.mm_b200_kh_replay.restype = C.c_uint32
    so.mm_b200_kh_replay.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    return so


def replay(so, hashes):
    h = np.ascontiguousarray(hashes, dtype=np.uint32)
    cap = 8
    while cap < 2 * len(h):
        cap *= 2
    out = np.empty(cap * 2, dtype=np.int32)
    nb = so.mm_b200_kh_replay(h.ctypes.data, len(h), out.ctypes.data)
    s = out[:nb]
    return nb, s[s >= 0]


@pytest.mark.parametrize("size,args", [(300_000, ["-k", "15", "-w", "5"]), (3_000_000, ["-k", "21", "-w", "11"]),
                                       (1_500_000, ["-H", "-k", "19", "-w", "10"])])
def test_replay_matches_reference_index_files(lib, tmp_path, size, args):
    if not os.path.exists(L.REF_BIN_B):
        pytest.skip("needs oracle/_ref/minimap2_B")
    rng = random.Random(size)
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as f:
        for c in range(3):
            f.write(f">c{c}\n" + "".join(rng.choice("ACGT") for _ in range(size // 3)) + "\n")
    subprocess.check_call([L.REF_BIN_B] + args + ["-d", str(tmp_path / "ref.mmi"), str(fa)], stderr=subprocess.DEVNULL)
    (part,) = parse_mmi(tmp_path / "ref.mmi")
    n_grown = 0
    for p, keys, vals in part["buckets"]:
        if len(keys) == 0:
            continue
        order = np.argsort(keys >> np.uint64(1), kind="stable")        # insertion order: ascending minimizer
        h = ((keys[order] >> np.uint64(1)) & np.uint64(0xffffffff)).astype(np.uint32)
        nb, got = replay(lib, h)
        assert len(got) == len(keys)
        assert np.array_equal(order[got], np.arange(len(keys))), "slot order differs from the reference's file"
        nb0 = 4
        while nb0 < len(keys):
            nb0 *= 2
        n_grown += nb != nb0
    assert n_grown > 100     # tables that doubled while being filled (the in-place rehash) were exercised


def test_replay_small_and_degenerate(lib):
    nb, got = replay(lib, [7])
    assert nb == 4 and list(got) == [0]
    nb, got = replay(lib, [0, 4, 8, 12])          # all collide modulo 4; the 4th insertion doubles the table
    assert nb == 8 and sorted(got) == [0, 1, 2, 3]
