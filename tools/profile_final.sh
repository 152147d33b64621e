# refresh of the round's profiling artefacts for the shipped callback kernel (one GPU, under gpurun)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 10 -c 16 --csv --log-file gpurun_out/$1_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scene_mix -s 3 -c 1 -f -o gpurun_out/$1_smx python bench.py --steps 4 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
ls -la gpurun_out/$1_*
