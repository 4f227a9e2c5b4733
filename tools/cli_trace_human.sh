#!/bin/bash
# development aid: where the whole CLI spends its mapping phase at human size (MM2_B200_TRACE=1 prints every stage with its wall time)
D=/dev/shm/airlift_b200_bench; mkdir -p $D
P=$D/pair_3100000000_24
[ -f $P.ok ] || { build/mmsynth pair $P 3100000000 24 42 43 2>/dev/null; touch $P.ok; }
[ -f $D/c_1.fq ] || build/mmsynth srp $P $D/c_1.fq $D/c_2.fq 3000000 45
MM2_B200_TRACE=1 build/minimap2-b200 -ax sr -t 16 -K 150000000 $P.new.fa $D/c_1.fq $D/c_2.fq > /dev/null 2> gpurun_out/cli_trace_human.err
grep -c . gpurun_out/cli_trace_human.err; grep "M::main\|worker_pipeline" gpurun_out/cli_trace_human.err | cut -c1-160
