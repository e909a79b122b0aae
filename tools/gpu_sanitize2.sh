#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernels added late in round 2: short-key attention, compile-time temporal
# attention, GroupNorm with cache hints (small and ragged shapes; the 6144-row cases are skipped: minutes under the tool)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export CCEDIT_CUDA_GRAPH=0
SEL='test_attention_short_keys or test_attention_many_heads or test_temporal_attention or test_groupnorm or test_attention_text'
DESEL='not 6144 and not 1536'
SAN="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600"
timeout 1500 $SAN python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "($SEL) and $DESEL" > gpurun_out/sanitize2_mem.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitize2_mem.log
SAN="compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600"
timeout 1500 $SAN python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "(test_attention_short_keys or test_temporal_attention) and $DESEL" > gpurun_out/sanitize2_race.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/sanitize2_race.log
CCEDIT_TA_FIXED=0 timeout 1500 $SAN python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "test_temporal_attention" > gpurun_out/sanitize2_race_generic.log 2>&1; echo "racecheck (generic temporal kernel) exit $?"; tail -4 gpurun_out/sanitize2_race_generic.log
timeout 300 python tools/dev_norm.py 2>&1 | grep "t_attn" | head -4
