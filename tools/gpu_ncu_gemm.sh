#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one_gemm.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_gemm as dg
from ccedit_b200 import ops
which = sys.argv[1]
if which == "conv":
    dg.run(0, 320, 320, taps=ops.conv_taps(), shape=(34, 64, 96), iters=2)
elif which == "lin640":
    dg.run(52224, 640, 640, res=True, iters=2)
elif which == "geglu":
    dg.run(208896, 320, 2560, geglu=True, iters=2)
PY
for w in conv lin640 geglu; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 1 -c 1 -f -o gpurun_out/r02_gemm_$w python /tmp/one_gemm.py $w > gpurun_out/ncu_gemm_$w.log 2>&1; echo "ncu $w exit $?"
done
CCEDIT_GEMM_CLUSTER=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 1 -c 1 -f -o gpurun_out/r02_gemm_conv_nocluster python /tmp/one_gemm.py conv > gpurun_out/ncu_gemm_conv_nc.log 2>&1; echo "ncu conv nocluster exit $?"
