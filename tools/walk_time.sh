#!/bin/bash
# prints the ncu duration of k_walk_seek on the C3 bench
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_walk" -c 6 --csv --log-file gpurun_out/walk_time.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep walk gpurun_out/walk_time.csv | tail -2 | awk -F'","' '{print "walk ns:", $NF}'
