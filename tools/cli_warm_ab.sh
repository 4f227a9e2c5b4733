#!/bin/bash
# development aid: mapping-phase seconds of the CLI at human size for several MM2_B200_WARM_BATCHES settings (needs tools/cli_trace_human.sh's files)
D=/dev/shm/airlift_b200_bench; P=$D/pair_3100000000_24
[ -f $D/c_1.fq ] || tools/cli_trace_human.sh > /dev/null
for rep in 1 2; do for w in 16 2 4; do
  MM2_B200_WARM_BATCHES=$w build/minimap2-b200 -ax sr -t 16 -K 150000000 $P.new.fa $D/c_1.fq $D/c_2.fq 2> /tmp/e.$w > /dev/null
  a=$(grep "loaded/built" /tmp/e.$w | sed 's/.*main::\([0-9.]*\)\*.*/\1/'); b=$(grep "Real time" /tmp/e.$w | sed 's/.*Real time: \([0-9.]*\) sec.*/\1/')
  l=$(grep "worker_pipeline" /tmp/e.$w | tail -1 | sed 's/.*pipeline::\([0-9.]*\)\*.*/\1/')
  echo "warm=$w rep=$rep index=$a last_mapped=$l end=$b"
done; done
