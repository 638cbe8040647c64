set +e
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2/smi.txt 2>&1
# 1. experimental fused Mlp kernel A: plain, then memcheck
DPMN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental.py -x -q > gpurun_out/r2/exp_a.log 2>&1
echo "exp_a rc=$?" >> gpurun_out/r2/exp_a.log
DPMN_EXPERIMENTAL=1 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2/exp_a_memcheck.log python -m pytest tests/test_experimental.py -x -q -k "1" > gpurun_out/r2/exp_a_memcheck.out 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2/exp_a_memcheck.out
# 2. racecheck / memcheck on the 4 mbarrier-pipelined kernels through small fp16 cases
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2/racecheck_tc.log python -m pytest tests/test_gpu_parity.py -x -q -k "test_pgrm_tensor_core_modes_match_reference and fp16 and pgrm_i0_m0 or test_cmm_tensor_core_modes_match_reference and fp16" > gpurun_out/r2/racecheck_tc.out 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2/racecheck_tc.out
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2/memcheck_tc.log python -m pytest tests/test_gpu_parity.py -x -q -k "test_pgrm_tensor_core_modes_match_reference and fp16 and pgrm_i0_m0 or test_cmm_tensor_core_modes_match_reference and fp16" > gpurun_out/r2/memcheck_tc.out 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2/memcheck_tc.out
# 3. baseline numbers of the round-1 state on this box
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_train_base.json 2> gpurun_out/r2/bench_train_base.err
tail -3 gpurun_out/r2/exp_a.log gpurun_out/r2/exp_a_memcheck.out gpurun_out/r2/racecheck_tc.out gpurun_out/r2/memcheck_tc.out
cat gpurun_out/r2/bench_base.json | cut -c1-400
cat gpurun_out/r2/bench_train_base.json | cut -c1-400
