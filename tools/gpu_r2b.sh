#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_boundary_gpu.py -m gpu -q -x > gpurun_out/pytest_net.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_net.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit $?"; tail -5 gpurun_out/bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['network_call'], d.get('variants'))
print(d['kernel_families'][:4])
for k in ('config4_per_gpu_share','config3_tvi2v'): print(k, d['configs'][k])
PY
