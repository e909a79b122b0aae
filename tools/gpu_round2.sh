#!/bin/bash
# GPU tests + ncu launch list + ncu full captures of the two top kernels + bench.  Run via gpurun from the repo root.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/one_call.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flash_attn -c 2 \
    -f -o gpurun_out/attn python tools/one_call.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tap_gemm -s 60 -c 4 \
    -f -o gpurun_out/gemm python tools/one_call.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
