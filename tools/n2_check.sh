set -e
D=/dev/shm/n2; mkdir -p $D
build/mmsynth ref $D/ref.fa 20000000 4 42
build/mmsynth sr $D/ref.fa $D/r1.fq $D/r2.fq 200000 44
build/minimap2-b200 -ax sr -t 16 --gpus 2 $D/ref.fa $D/r1.fq $D/r2.fq 2> $D/new.err | grep -v '^@PG' > $D/new.sam
oracle/_ref/minimap2_B -ax sr -t 16 $D/ref.fa $D/r1.fq $D/r2.fq 2> $D/ref.err | grep -v '^@PG' > $D/ref.sam
cmp $D/new.sam $D/ref.sam && echo "IDENTICAL with --gpus 2: $(wc -l < $D/ref.sam) lines"
tail -2 $D/new.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"per_rank_ms_per_step": \[[0-9., ]*\]' gpurun_out/bench_n2c.json | head -6
