#!/bin/bash
# GEMM epilogue with pipelined TMEM loads: parity suites, the shape timings of tools/dev_gemm.py, a short bench.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/dev_gemm.py > gpurun_out/dev_gemm_pipe.txt 2>&1
cp ccedit_b200/lib/libccedit_b200.so /tmp/pipe.so; cp tools/_bin/libccedit_nopipe.so ccedit_b200/lib/libccedit_b200.so
timeout 600 python tools/dev_gemm.py > gpurun_out/dev_gemm_nopipe.txt 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-configs > gpurun_out/bench_noepi.json 2> gpurun_out/bench_noepi.err; echo "bench(nopipe) exit $?"
cp /tmp/pipe.so ccedit_b200/lib/libccedit_b200.so
paste -d'\n' gpurun_out/dev_gemm_pipe.txt gpurun_out/dev_gemm_nopipe.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-configs > gpurun_out/bench_epi.json 2> gpurun_out/bench_epi.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_noepi.json','gpurun_out/bench_epi.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['network_call']['ms'], d['clocks'])
    for k in d['kernels'][:6]: print(k)
PY
