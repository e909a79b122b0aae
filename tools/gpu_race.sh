#!/bin/bash
# racecheck (shared-memory hazards) over the kernels that synchronise with __syncthreads / __syncwarp rather than mbarriers:
# GroupNorm, LayerNorm, hint stem, layout + small kernels, CLIP helpers, first-stage softmax; and, for information, the
# tcgen05 kernels (the tool does not model mbarrier / TMA ordering: hazards there need reading, not counting)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export CCEDIT_CUDA_GRAPH=0
SAN="compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600"
timeout 1500 $SAN python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_groupnorm or test_layernorm or test_hint_stem or test_layout" > gpurun_out/race_plain.log 2>&1; echo "racecheck plain kernels exit $?"; tail -3 gpurun_out/race_plain.log
timeout 1500 $SAN python -m pytest tests/test_vae_gpu.py -m gpu -q -k "clip or attn_512 and not big" > gpurun_out/race_vae.log 2>&1; echo "racecheck clip / vae attention exit $?"; tail -3 gpurun_out/race_vae.log
timeout 1500 $SAN python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_linear and not staged or test_attention and not short and not 6144 and not 1536" > gpurun_out/race_tc.log 2>&1; echo "racecheck tcgen05 kernels exit $?"; tail -3 gpurun_out/race_tc.log; grep -c "Error: Race" gpurun_out/race_tc.log
