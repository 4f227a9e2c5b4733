set -e
D=/dev/shm/prof; mkdir -p $D
build/mmsynth ref $D/ref.fa 100000000 6 42
build/mmsynth sr $D/ref.fa $D/r1.fq $D/r2.fq 2000000 44
( time build/minimap2-b200 -ax sr -t 16 -K 150M $D/ref.fa $D/r1.fq $D/r2.fq > /dev/null ) 2> gpurun_out/cli_time.err
LD_PRELOAD=build/sampler.so SAMPLER_OUT=gpurun_out/samp_cli.out build/minimap2-b200 -ax sr -t 16 -K 150M $D/ref.fa $D/r1.fq $D/r2.fq > /dev/null 2> gpurun_out/samp_cli.err
tail -25 gpurun_out/cli_time.err
( time oracle/_ref/minimap2_B -ax sr -t 16 $D/ref.fa $D/r1.fq $D/r2.fq > /dev/null ) 2>&1 | tail -6
