#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
echo "== dev_gemm CLUSTER=4"; CCEDIT_GEMM_CLUSTER=4 timeout 300 python tools/dev_gemm.py 2>&1 | tail -11
echo "== kernels CLUSTER=4"; CCEDIT_GEMM_CLUSTER=4 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-configs > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench exit $?"; tail -3 gpurun_out/bench_pair.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_pair.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['network_call'])
for k in d['kernels'][:8]: print(k)
PY
CCEDIT_GEMM_CLUSTER=1 timeout 900 python bench.py --no-cpu-baseline --no-configs --no-breakdown > gpurun_out/bench_mc.json 2> gpurun_out/bench_mc.err; echo "bench mc exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_mc.json'))
print('multicast:', d['value'], d['ms_per_step'])
PY
