# development aid: time the CLI (with per-batch trace) next to the reference fork on the same files
set -e
D=/dev/shm/prof; mkdir -p $D
[ -f $D/ref.fa ] || build/mmsynth ref $D/ref.fa 100000000 6 42
[ -f $D/r1.fq ] || build/mmsynth sr $D/ref.fa $D/r1.fq $D/r2.fq 2000000 44
MM2_B200_TRACE=1 build/minimap2-b200 -ax sr -t 16 -K 150M $D/ref.fa $D/r1.fq $D/r2.fq 2> gpurun_out/cli_trace.err > /dev/null
grep "T::\|M::main\|M::worker" gpurun_out/cli_trace.err
