# warp-specialised callback kernel on the GPU box: parity subset, then bench lines per ODB_SMX_CFG
#   tools/exp_ws.sh <tag> <cfg> [cfg ...]       (a trailing "ncu" captures the last successful cfg with ncu --set full)
tag=$1; shift
last_ok=""
for c in "$@"; do
  if [ "$c" = "ncu" ]; then
    [ -n "$last_ok" ] && ODB_SMX_CFG=$last_ok timeout 90 ncu --set full --clock-control none --import-source on -k regex:k_scene_mix -s 3 -c 1 -f -o gpurun_out/${tag}_smx_cfg$last_ok python bench.py --steps 4 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
    continue
  fi
  if [ "$c" != "0" ]; then
    ODB_SMX_CFG=$c timeout 45 python -m pytest tests/test_scene_gpu.py tests/test_golden.py tests/test_cycle_gpu.py -x -q -m gpu > gpurun_out/${tag}_pytest_cfg$c.log 2>&1
    rc=$?
    echo "cfg $c pytest rc=$rc"; tail -3 gpurun_out/${tag}_pytest_cfg$c.log
    [ $rc -ne 0 ] && continue
  fi
  ODB_SMX_CFG=$c timeout 45 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/${tag}_cfg$c.json 2> gpurun_out/${tag}_cfg$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_cfg$c.json"))
    print("cfg $c: step %.1f us kernel %.1f us frac %.3f checksum %.9f" % (d["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["checksum"]))
except Exception as e:
    print("cfg $c: FAILED", e)
PY
  [ -s gpurun_out/${tag}_cfg$c.json ] && last_ok=$c
done
