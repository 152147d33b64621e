#!/bin/bash
# Runs bench.py for several kernel variants and prints value / step / kernel time for each.
mkdir -p gpurun_out
for v in "$@"; do
  python bench.py --no-cpu-baseline --variant $v > gpurun_out/bv_$v.json 2> gpurun_out/bv_$v.err || tail -3 gpurun_out/bv_$v.err
  python -c "import json;d=json.load(open('gpurun_out/bv_$v.json'));print('variant $v', 'value %.4g' % d['value'], 'step_us %.1f' % (d['ms_per_step']*1e3), 'kernel_us %.1f' % (d['roofline']['kernel_ms']*1e3), 'frac %.3f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'])"
done
python -c "import json,sys;d=json.load(open(\"gpurun_out/bv_$1.json\"));print(\"host enqueue us/step\", d[\"config\"][\"host_enqueue_us_per_step\"])"
