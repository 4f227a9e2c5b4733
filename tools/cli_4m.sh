# development aid: the CLI run bench.py times (4 M pairs, -K 150M), with the per-batch trace
set -e
D=/dev/shm/prof4; mkdir -p $D
[ -f $D/ref.fa ] || build/mmsynth ref $D/ref.fa 100000000 6 42
[ -f $D/c1.fq ] || build/mmsynth sr $D/ref.fa $D/c1.fq $D/c2.fq 4000000 45
MM2_B200_TRACE=1 build/minimap2-b200 -ax sr -t $(nproc) -K 150M $D/ref.fa $D/c1.fq $D/c2.fq 2> gpurun_out/cli4m.err > /dev/null
grep "T::map_step\|T::read\|T::write\|M::main\|M::worker" gpurun_out/cli4m.err | cut -c1-220
