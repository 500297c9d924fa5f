#!/bin/bash
# compute-sanitizer + long fuzz over the GPU suite (through gpurun) -> gpurun_out/<r>_sanitizer.txt     usage: bash scripts/sanitize.sh r02
R=${1:-r02}
mkdir -p gpurun_out
OUT=gpurun_out/${R}_sanitizer.txt
echo "# compute-sanitizer (CUDA 12.9) over the GPU test-suite on a B200, final kernels of round 2" > $OUT
echo '$ compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "not full_size and not 2_31 and not larger_than and not at_scale and not plugin and not opt_stack and not parallel"' >> $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "not full_size and not 2_31 and not larger_than and not at_scale and not plugin and not opt_stack and not parallel" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | tail -6 >> $OUT
echo '$ compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_softmax_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "softmax_equals_torch_bitwise or post_chain or mask_add or histc_vs or minmax or histogram_observer or tma"   # the kernels that use shared memory' >> $OUT
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_softmax_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "softmax_equals_torch_bitwise or post_chain or mask_add or histc_vs or minmax or histogram_observer or tma" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -6 >> $OUT
echo '$ DMXQ_FUZZ_SEEDS=300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k fuzz' >> $OUT
DMXQ_FUZZ_SEEDS=300 timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -k fuzz 2>&1 | grep -E "passed|failed" | tail -3 >> $OUT
cat $OUT
