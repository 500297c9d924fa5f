OUT=gpurun_out/r02_sanitizer_final.txt
echo "# final tree of round 2 (after the softmax code-size refactor, the K_FLOAT / K_BFP_STOCH extensions, the bias cache)" > $OUT
echo '$ compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "not full_size and not 2_31 and not larger_than and not at_scale and not plugin and not opt_stack and not parallel"' >> $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "not full_size and not 2_31 and not larger_than and not at_scale and not plugin and not opt_stack and not parallel" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | tail -4 >> $OUT
echo '$ compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_softmax_gpu.py -m gpu -q -x -k "not at_scale and (bitwise or mask_add or nan_and_inf)"' >> $OUT
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_softmax_gpu.py -m gpu -q -x -k "not at_scale and (bitwise or mask_add or nan_and_inf)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -4 >> $OUT
cat $OUT
