#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py -m gpu -q > gpurun_out/pytest_vae.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_vae.log
timeout 300 python tools/dev_attn.py 2>&1 | grep -E "BAD|attn F=34|cross-attn"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?"; tail -5 gpurun_out/bench3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench3.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['network_call'])
print(d['configs']['first_stage_decode'])
for k in d['kernels'][:8]: print(k)
PY
