# development aid: long-read throughput over several mini-batches next to the reference fork
set -e
D=/dev/shm/ont2; mkdir -p $D
N=${1:-60000}
[ -f $D/ref.fa ] || build/mmsynth ref $D/ref.fa 100000000 6 42
[ -f $D/long.fq ] || build/mmsynth long $D/ref.fa $D/long.fq $N 45
MM2_B200_TRACE=1 build/minimap2-b200 -ax map-ont -t 16 -K 200M $D/ref.fa $D/long.fq 2> gpurun_out/ont2_new.err | grep -v '^@PG' > $D/new.sam
oracle/_ref/minimap2_B -ax map-ont -t 16 -K 200M $D/ref.fa $D/long.fq 2> gpurun_out/ont2_ref.err | grep -v '^@PG' > $D/ref.sam
grep "T::map_step\|Real time\|loaded" gpurun_out/ont2_new.err | cut -c1-250
grep "Real time\|loaded" gpurun_out/ont2_ref.err
cmp $D/new.sam $D/ref.sam && echo "IDENTICAL $(wc -l < $D/ref.sam) lines"
