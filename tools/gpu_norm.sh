#!/bin/bash
# GroupNorm / temporal attention at the headline network's shapes: isolated timings over rotating buffers (tools/dev_norm.py)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python tools/dev_norm.py 2>&1 | tee gpurun_out/dev_norm.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "norm or gemm" 2>&1 | tail -3
