#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['network_call']['ms'], d['clocks'])
for k in d['kernels'][:6]: print(k)
PY
