# end-of-round check on the GPU box (one GPU): the whole GPU suite on the shipped default, then the default bench line
timeout 150 python -m pytest tests -q -m gpu -s > gpurun_out/$1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/$1_pytest_gpu.log
timeout 200 python bench.py > gpurun_out/$1_bench_n1.json 2> gpurun_out/$1_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/$1_bench_n1.json"))
print("step %.1f us kernel %.1f us frac %.3f e2e %.3g checksum %.9f" % (d["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["e2e"]["value"], d["checksum"]))
PY
timeout 60 python bench.py --variant 0 --no-cpu-baseline --skip-e2e > gpurun_out/$1_bench_n1_strict.json 2>> gpurun_out/$1_bench_n1.err; echo "strict rc=$?"
