set -x
cd /root/repo
# launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_launches.log 2>&1
# full captures of the dominant kernel, both variants
ncu --set full --clock-control none --import-source on -k regex:k_mix_fast -s 3 -c 1 -o gpurun_out/r1b_fast_fma python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mix_fast -s 3 -c 1 -o gpurun_out/r1b_fast_strict python bench.py --steps 4 --warmup 3 --no-cpu-baseline --variant 0 > /dev/null 2>&1
# bench lines (not under a profiler)
python bench.py > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err
python bench.py --variant 0 --no-cpu-baseline > gpurun_out/r1b_bench_n1_strict.json 2>> gpurun_out/r1b_bench_n1.err
python bench.py --impl reference > gpurun_out/r1b_bench_reference_arm.json 2>> gpurun_out/r1b_bench_n1.err
python tools/bench_mixer.py > gpurun_out/r1b_c4.log 2>&1
python tools/bench_buffered.py > gpurun_out/r1b_c3b.log 2>&1
tail -2 gpurun_out/r1b_c4.log gpurun_out/r1b_c3b.log
