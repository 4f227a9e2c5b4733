#!/bin/bash
# distribution of the data-dependent work per pass on the human-size workload (debug aid): MMG_TRACE prints one line per pass
set -e
D=/dev/shm/airlift_b200_bench; mkdir -p $D
P=$D/pair_3100000000_24
[ -f $P.ok ] || { build/mmsynth pair $P 3100000000 24 42 43 2>/dev/null; touch $P.ok; }
build/mmsynth srp $P $D/t_1.fq $D/t_2.fq ${1:-500000} 44
MMG_TRACE=1 MM2_B200_PROFILE=1 build/minimap2-b200 -ax sr -t 16 -K 150000000 $P.new.fa $D/t_1.fq $D/t_2.fq 2> gpurun_out/trace_pass.err > /dev/null
grep "mmg::pass\|paths\|kernel" gpurun_out/trace_pass.err | head -80
