#!/bin/bash
# compute-sanitizer memcheck over the kernels changed in round 2 (small shapes): pair-MMA / multicast GEMM incl. phantom
# tiles, staged epilogue with warp-collective TMA, attention with barrier probes, sampler kernels, dup_rows, VAE helpers.
mkdir -p gpurun_out
export CCEDIT_CUDA_GRAPH=0
SAN="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600"
timeout 1500 $SAN python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "test_linear or test_conv3x3 or test_temporal_conv or test_attention or test_layout" > gpurun_out/sanitize_kernels.log 2>&1; echo "memcheck kernels exit $?"; tail -6 gpurun_out/sanitize_kernels.log
timeout 1200 $SAN python -m pytest tests/test_vae_gpu.py -m gpu -q -x -k "block or decode_video or encode" > gpurun_out/sanitize_vae.log 2>&1; echo "memcheck vae exit $?"; tail -4 gpurun_out/sanitize_vae.log
timeout 1200 $SAN python -m pytest tests/test_network_gpu.py -m gpu -q -x -k "fused_sampler_is_bit_identical_to_unfused and False or forward_cfg" > gpurun_out/sanitize_net.log 2>&1; echo "memcheck net exit $?"; tail -4 gpurun_out/sanitize_net.log
for m in 2 4; do CCEDIT_GEMM_CLUSTER=$m timeout 900 $SAN python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "test_linear or test_conv3x3" > gpurun_out/sanitize_cl$m.log 2>&1; echo "memcheck cluster=$m exit $?"; tail -3 gpurun_out/sanitize_cl$m.log; done
grep -h "ERROR SUMMARY" gpurun_out/sanitize_*.log | sort | uniq -c
