#!/bin/bash
# programmatic dependent launch (CCEDIT_PDL bit mask: 1 tap-GEMM, 2 flash attention, 4 short-key / temporal attention,
# 8 GroupNorm) against plain stream order: bench A/B on one box
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
for m in ${MASKS:-0 1 3 5 9 0 1}; do
  CCEDIT_PDL=$m timeout 600 python bench.py --steps 8 --warmup 3 --no-configs > gpurun_out/bench_pdl$m.json 2> gpurun_out/bench_pdl$m.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl$m.json').read().strip().splitlines()[-1])
print('PDL', $m, d['value'], d['ms_per_step'], d['e2e']['value'], d['network_call']['ms'], d['clocks']['sm_mhz'])
PY
done
