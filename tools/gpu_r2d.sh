#!/bin/bash
# Round 2 evidence pass: all GPU tests, smoke, parity report, bench (both arms), ncu launch list with DRAM traffic.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python tools/parity_report.py > gpurun_out/parity.md 2> gpurun_out/parity.err; echo "parity exit $?"; tail -16 gpurun_out/parity.md; tail -3 gpurun_out/parity.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['network_call'], d.get('variants'), d.get('library_baseline'), d.get('cpu_baseline',{}).get('value'))
print(d['roofline'])
for k in d['kernels'][:10]: print(k)
print({k:(v if not isinstance(v,list) else '...') for k,v in d['configs'].items()})
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/one_call.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"; tail -2 gpurun_out/ncu_traffic.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cut -c1-700 gpurun_out/bench_ref.json
