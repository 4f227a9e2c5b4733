#!/bin/bash
# development aid: A/B runs of the short-read bench under the kernels' environment switches (one gpurun call)
B="python bench.py --no-cli --no-cpu-baseline --no-long-read --no-parity --steps 4 --warmup 3"
$B > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
MMG_TAIL_HEAVY_N=1024 $B > gpurun_out/ab_heavy1024.json 2> gpurun_out/ab_heavy1024.err
MMG_TAIL_HEAVY_N=4096 $B > gpurun_out/ab_heavy4096.json 2> gpurun_out/ab_heavy4096.err
MMG_TAIL_WALK=1 $B > gpurun_out/ab_walk.json 2> gpurun_out/ab_walk.err
echo done
