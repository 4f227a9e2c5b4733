#!/usr/bin/env python
"""top source lines by warp-stall samples of one kernel launch in an ncu report (compiled with -lineinfo, captured with --import-source on)
usage: ncu_lines.py report.ncu-rep <kernel-id e.g. ::k_heap_replay:2> [n]"""
import csv, subprocess, sys, io
csv.field_size_limit(10**9)
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", kid], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0]:
        d = dict(zip(hdr[4:], r[4:]))
        stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        lines.append((int(d["# Samples"] or 0), cur_file, r[0], r[1].strip()[:110], d["Thread Instructions Executed"], d["Avg. Threads Executed"], stalls))
tot = sum(l[0] for l in lines)
print("total samples", tot)
for s, f, ln, src, ti, at, st in sorted(lines, key=lambda l: -l[0])[:n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100.0 * s / max(tot,1):5.1f}% {f}:{ln:>5} thr_inst={ti:>10} avg_thr={at:>3} {src}\n        {top}")
