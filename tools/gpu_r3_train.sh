set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train_tail.py tests/test_gpu_edges.py -q -s > $O/gputests_train.log 2>&1
echo "rc=$?" >> $O/gputests_train.log
grep -E "pgrm fp16|pgrm bf16|passed|failed|rc=|Error|assert" $O/gputests_train.log | tail -15
