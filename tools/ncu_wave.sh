D=/dev/shm/ont; mkdir -p $D
[ -f $D/ref.fa ] || build/mmsynth ref $D/ref.fa 100000000 6 42
[ -f $D/long.fq ] || build/mmsynth long $D/ref.fa $D/long.fq 20000 45
ncu --set full --clock-control none --import-source on -k k_ksw_wave -c 2 -o /tmp/prof_wave -f build/minimap2-b200 -ax map-ont -t 16 $D/ref.fa $D/long.fq > /dev/null 2> gpurun_out/ncu_wave.log
ncu -i /tmp/prof_wave.ncu-rep --page raw --csv > gpurun_out/ncu_wave_raw.csv 2>/dev/null
ncu -i /tmp/prof_wave.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > gpurun_out/ncu_wave_src.csv 2>/dev/null
ls -la gpurun_out/ncu_wave*
