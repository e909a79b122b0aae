#!/bin/bash
mkdir -p gpurun_out
echo "== pytest kernels CLUSTER=3"; CCEDIT_GEMM_CLUSTER=3 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -m gpu -q -x > gpurun_out/pytest_cl3.log 2>&1; echo "exit $?"; tail -4 gpurun_out/pytest_cl3.log
for m in 3 1; do echo "== dev_gemm CCEDIT_GEMM_CLUSTER=$m"; CCEDIT_GEMM_CLUSTER=$m timeout 300 python tools/dev_gemm.py 2>&1 | tail -11; done
cat > /tmp/one_gemm.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_gemm as dg
from ccedit_b200 import ops
dg.run(0, 320, 320, taps=ops.conv_taps(), shape=(34, 64, 96), iters=2)
PY
CCEDIT_GEMM_CLUSTER=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 1 -c 1 -f -o gpurun_out/r02_gemm_conv_pair python /tmp/one_gemm.py > gpurun_out/ncu_gemm_conv_pair.log 2>&1; echo "ncu pair exit $?"
