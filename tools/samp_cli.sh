D=/dev/shm/prof; mkdir -p $D
[ -f $D/ref.fa ] || build/mmsynth ref $D/ref.fa 100000000 6 42
[ -f $D/r1.fq ] || build/mmsynth sr $D/ref.fa $D/r1.fq $D/r2.fq 2000000 44
LD_PRELOAD=build/sampler.so SAMPLER_OUT=gpurun_out/samp_cli2.out build/minimap2-b200 -ax sr -t 16 -K 150M $D/ref.fa $D/r1.fq $D/r2.fq > /dev/null 2> gpurun_out/samp_cli2.err
nm -D --defined-only /lib/x86_64-linux-gnu/libc.so.6 | sort > gpurun_out/libc.syms
wc -l gpurun_out/samp_cli2.out
