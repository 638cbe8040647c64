set +e
O=gpurun_out/r3; mkdir -p $O
timeout 1500 compute-sanitizer --tool initcheck --print-limit 40 --log-file $O/initcheck_train_step.log python -m pytest tests/test_gpu_train_tail.py -m gpu -x -q -k "two_stream and fp16" > $O/initcheck_train_step.out 2>&1
echo "initcheck rc=$?"; tail -3 $O/initcheck_train_step.out
grep -c "Uninitialized" $O/initcheck_train_step.log
grep -A3 "Uninitialized" $O/initcheck_train_step.log | grep "at \|by thread" | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30
tail -3 $O/initcheck_train_step.log
