#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['network_call'], d['roofline'])
for k in d['kernels'][:12]: print(k)
PY
tail -3 gpurun_out/bench.err
timeout 300 python tools/gemm_shapes.py tv2v > gpurun_out/shapes_tv2v.txt 2>&1; head -40 gpurun_out/shapes_tv2v.txt
