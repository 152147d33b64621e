# kernel-shape experiments on the GPU box: tools/exp_cfg.sh <tag> <cfg> [cfg ...]  (ODB_SMX_CFG values; "L" = round 1's multi-kernel path)
tag=$1; shift
for c in "$@"; do
  if [ "$c" = "L" ]; then extra="--variant 0x202"; env_c=0; else extra=""; env_c=$c; fi
  ODB_SMX_CFG=$env_c python bench.py --steps 16 --warmup 3 --no-cpu-baseline --skip-e2e $extra > gpurun_out/${tag}_cfg$c.json 2> gpurun_out/${tag}_cfg$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_cfg$c.json"))
    print("cfg $c: step %.1f us kernel %.1f us frac %.3f checksum %.6f" % (d["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["checksum"]))
except Exception as e:
    print("cfg $c: FAILED", e)
PY
done
