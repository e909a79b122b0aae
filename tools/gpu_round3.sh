#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -s 11 -c 1 \
    -f -o gpurun_out/attn_tc python tools/dev_attn.py > gpurun_out/ncu_attn_tc.log 2>&1; echo "ncu attn exit $?"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['network_call'], d['roofline'])
for k in d['kernels'][:12]: print(k)
PY
tail -5 gpurun_out/bench.err
