set +e
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train_tail.py tests/test_gpu_edges.py -q -s -k "same_arithmetic or train_tail or fused or allreduce or prepared or alpha or single_32x32 or error_behaviour" 2>&1 | grep -E "passed|failed|worst|Error|assert " | cut -c1-400 | tail -n 30
