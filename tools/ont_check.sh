# development aid: long-read (map-ont) run of the CLI next to the reference fork, outputs compared
set -e
D=/dev/shm/ont; mkdir -p $D
N=${1:-20000}
[ -f $D/ref.fa ] || build/mmsynth ref $D/ref.fa 100000000 6 42
[ -f $D/long.fq ] || build/mmsynth long $D/ref.fa $D/long.fq $N 45
MM2_B200_PROFILE=1 MM2_B200_TRACE=1 MMG_KSW_DEBUG=1 build/minimap2-b200 -ax map-ont -t 16 $D/ref.fa $D/long.fq 2> gpurun_out/ont_new.err | grep -v '^@PG' > $D/new.sam
oracle/_ref/minimap2_B -ax map-ont -t 16 $D/ref.fa $D/long.fq 2> gpurun_out/ont_ref.err | grep -v '^@PG' > $D/ref.sam
grep "T::\|M::b200\|mmg_ksw\|Real time" gpurun_out/ont_new.err | cut -c1-400
grep "Real time\|loaded" gpurun_out/ont_ref.err
cmp $D/new.sam $D/ref.sam && echo "IDENTICAL $(wc -l < $D/ref.sam) lines"
