#!/bin/bash
# Round-end measurement set: tests, launch list with DRAM traffic, full ncu captures of the top kernels, benches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/one_call.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc2 -s 21 -c 1 \
    -f -o gpurun_out/final_attn python tools/dev_attn.py > gpurun_out/ncu_final_attn.log 2>&1; echo "ncu attn exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 48 -c 1 \
    -f -o gpurun_out/final_gemm_conv python tools/dev_gemm.py > gpurun_out/ncu_final_gemm.log 2>&1; echo "ncu gemm conv exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 13 -c 1 \
    -f -o gpurun_out/final_gemm_geglu python tools/dev_gemm.py > gpurun_out/ncu_final_gemm2.log 2>&1; echo "ncu gemm geglu exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 1 -c 1 \
    -f -o gpurun_out/final_gemm_staged python tools/dev_gemm.py > gpurun_out/ncu_final_gemm3.log 2>&1; echo "ncu gemm staged exit $?"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --kind tvi2v --no-cpu-baseline --steps 5 > gpurun_out/bench_tvi2v.json 2> gpurun_out/bench_tvi2v.err; echo "bench tvi2v exit $?"; cut -c1-300 gpurun_out/bench_tvi2v.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cut -c1-500 gpurun_out/bench_ref.json
