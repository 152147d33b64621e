# Re-creates the round's profiling artefacts on the GPU box (run under gpurun, one GPU): launch list, full ncu captures of
# the two kernels of a C3 callback, the bench lines. Summaries for profiles/ are made here afterwards with
# tools/ncu_summary.py, tools/ncu_regions.py and tools/sass_excerpt.py.
set -x
cd /root/repo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 10 -c 16 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scene_mix -s 3 -c 1 -o gpurun_out/smx python bench.py --steps 4 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_walk_seek -s 3 -c 1 -o gpurun_out/walk python bench.py --steps 4 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
# bench lines (not under a profiler)
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --variant 0 --no-cpu-baseline --skip-e2e > gpurun_out/bench_n1_strict.json 2>> gpurun_out/bench_n1.err
python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2>> gpurun_out/bench_n1.err
for c in C1 C2 C3b C4 C5; do python bench.py --config $c --steps 8 > gpurun_out/bench_config_$c.json 2>> gpurun_out/bench_n1.err; done
