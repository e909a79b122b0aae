#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_shapes.py tv2v > gpurun_out/shapes_tv2v.txt 2>&1; echo "shapes exit $?"
timeout 200 python tools/dev_gemm.py > gpurun_out/dev_gemm.txt 2>&1; echo "dev_gemm exit $?"
CCEDIT_GEMM_TRACE=1 timeout 200 python tools/dev_gemm.py > gpurun_out/dev_gemm_trace.txt 2>&1; echo "trace exit $?"
head -100 gpurun_out/shapes_tv2v.txt
cat gpurun_out/dev_gemm.txt
