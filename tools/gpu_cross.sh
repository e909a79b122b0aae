#!/bin/bash
# text cross-attention (<= 128 keys shared by the frames of a batch entry): short-key kernel vs the long-sequence tcgen05
# path (CCEDIT_ATTN_SHORT=0), parity suites, a short bench with both
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py tests/test_network_gpu.py tests/test_boundary_gpu.py -x -q -m gpu 2>&1 | tail -3
cat > /tmp/cx.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_attn as da
for (L, d) in ((6144, 40), (1536, 80), (384, 160), (96, 160)):
    da.bench_cross(34, L, 77, 8, d)
PY
for m in 1 0; do echo "== SHORT=$m"; CCEDIT_ATTN_SHORT=$m timeout 300 python /tmp/cx.py 2>&1 | grep cross; done | tee gpurun_out/cross.txt
for m in 0 1; do
  CCEDIT_ATTN_SHORT=$m timeout 600 python bench.py --steps 8 --warmup 3 --no-configs > gpurun_out/bench_short$m.json 2> gpurun_out/bench_short$m.err; echo "bench SHORT=$m exit $?"
done
python - <<'PY'
import json
for m in (0, 1):
    d=json.loads(open(f'gpurun_out/bench_short{m}.json').read().strip().splitlines()[-1])
    print('SHORT', m, d['value'], d['ms_per_step'], d['network_call']['ms'], d['clocks'])
    for k in d['kernels'][:2]: print(k)
PY
