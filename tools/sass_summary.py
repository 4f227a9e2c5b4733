#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of the repo's K4 / K1 / K3 kernels (cuobjdump -sass on the objects `make lib` leaves in build/),
plus the instruction window around the first packed 16x2 instruction of k_ksw_dpx: the artefact behind DESIGN.md's statement
of which integer-SIMD (DPX) and asynchronous-copy instructions the hot kernels use.
usage: python tools/sass_summary.py > profiles/sass_r02.txt"""
import collections, re, subprocess, sys
KERNELS = ["k_ksw_dpx", "k_ksw_dpx_block", "k_ksw_tpj", "k_ksw_wave", "k_sketch_warp", "k_heap_replay", "k_chain_fill", "k_chain_fill_small",
           "k_chain_tail_block", "k_fill_flat_warp", "k_lookup"]
for obj in ["build/mmg_ksw.o", "build/mmg_stages.o"]:
    txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE).stdout.decode()
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name = part.split("\n", 1)[0].strip()
        short = next((k for k in KERNELS if name.startswith("_Z%d%s" % (len(k), k))), None)
        if short is None:
            continue
        lines = re.findall(r"/\*([0-9a-f]{4,6})\*/\s+((?:@!?U?P\d+\s+)?[A-Z][A-Za-z0-9_.]+[^;]*);", part)
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", l[1]).split()[0] for l in lines)
        tot = sum(ops.values())
        keep = {k: v for k, v in ops.items() if re.search(r"16x2|^VI|^PRMT|^REDUX|^LDGSTS|^SHFL|^VOTE|^ATOM|^RED|^LDS|^STS|^BAR|^MATCH", k)}
        print(f"== {short}  ({obj}, {tot} SASS instructions)")
        print("   " + ", ".join(f"{k} {v}" for k, v in sorted(keep.items(), key=lambda kv: -kv[1])))
        if short == "k_ksw_dpx":
            idx = next((i for i, l in enumerate(lines) if "16x2" in l[1] or ".S16" in l[1]), None)
            if idx is not None:
                print("   -- window around the first packed instruction (the cell-pair arithmetic of mmg_kswdpx_pair):")
                for a, t in lines[max(0, idx - 4): idx + 44]:
                    print(f"      /*{a}*/ {t.strip()}")
