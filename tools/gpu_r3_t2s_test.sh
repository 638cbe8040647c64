set +e
O=gpurun_out/r3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_train_tail.py -m gpu -x -q -s -k two_stream > $O/gputests_t2s_b.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_t2s_b.log
grep -n "fp16:\|fp32:\|passed\|failed\|rc=" $O/gputests_t2s_b.log
