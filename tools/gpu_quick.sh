#!/bin/bash
# quick check of a kernel change: kernel + block + network parity suites, then a short bench with the per-class table
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 8 --warmup 3 --no-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['network_call']['ms'], d['clocks'])
for k in d['kernels'][:10]: print(k)
PY
