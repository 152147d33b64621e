# kernel experiments on the GPU box: tools/exp.sh <tag> [lib ...]  - one bench line per library (device passes only)
tag=$1; shift
for so in "$@"; do
  name=$(basename $so .so)
  ODB_SO=$PWD/$so python bench.py --steps 16 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}: step %.1f us kernel %.1f us frac %.3f checksum %.6f" % (d["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["checksum"]))
except Exception as e:
    print("${name}: FAILED", e)
PY
done
