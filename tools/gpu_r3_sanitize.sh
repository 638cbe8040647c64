set +e
O=gpurun_out/r3; mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --log-file $O/memcheck_train_step.log python -m pytest tests/test_gpu_train_tail.py -m gpu -x -q -k "two_stream and fp16" > $O/memcheck_train_step.out 2>&1
echo "memcheck rc=$?"; tail -3 $O/memcheck_train_step.out; tail -4 $O/memcheck_train_step.log
timeout 1200 compute-sanitizer --tool racecheck --log-file $O/racecheck_train_step.log python -m pytest tests/test_gpu_train_tail.py -m gpu -x -q -k "two_stream and fp16" > $O/racecheck_train_step.out 2>&1
echo "racecheck rc=$?"; tail -3 $O/racecheck_train_step.out; tail -4 $O/racecheck_train_step.log
