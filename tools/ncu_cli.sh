#!/bin/bash
# ncu --set full capture (with source) of the kernels matching $1 in one human-size CLI run; report lands in gpurun_out/$2.ncu-rep
D=/dev/shm/airlift_b200_bench; mkdir -p $D
P=$D/pair_3100000000_24
[ -f $P.ok ] || { build/mmsynth pair $P 3100000000 24 42 43 2>/dev/null; touch $P.ok; }
[ -f $D/t_1.fq ] || build/mmsynth srp $P $D/t_1.fq $D/t_2.fq 500000 44
ncu --set full --clock-control none --import-source on -k "regex:$1" -c ${3:-8} -o gpurun_out/$2 -f \
  build/minimap2-b200 -ax sr -t 16 -K 150000000 $P.new.fa $D/t_1.fq $D/t_2.fq > /dev/null 2> gpurun_out/$2.log
tail -3 gpurun_out/$2.log; ls -la gpurun_out/$2.ncu-rep
# gpurun copies back at most 64 MiB: a larger report would cost the whole call's output
[ $(stat -c %s gpurun_out/$2.ncu-rep) -gt 60000000 ] && { echo "report too large, dropped"; rm -f gpurun_out/$2.ncu-rep; }
