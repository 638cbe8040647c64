set +e
O=gpurun_out/r3
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backward.py -q -s -k "attention_backward_tcgen05" > $O/bwd_tc_test.log 2>&1
echo "rc=$?" >> $O/bwd_tc_test.log
grep -E "passed|failed|rc=|Error|assert|dq|dkv|dtable" $O/bwd_tc_test.log | tail -20
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train_tail.py -q -s > $O/gputests_train2.log 2>&1
echo "rc=$?" >> $O/gputests_train2.log
grep -E "pgrm fp16|pgrm bf16|passed|failed|rc=|Error|assert" $O/gputests_train2.log | tail -12
DPMN_TRAIN_ATTN_BWD_TC=0 timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train_bwdsimt.json 2> $O/bench_train_bwdsimt.err
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_train_bwdtc.json 2> $O/bench_train_bwdtc.err
cut -c1-200 $O/bench_train_bwdsimt.json; cut -c1-200 $O/bench_train_bwdtc.json; tail -2 $O/bench_train_bwdtc.err
