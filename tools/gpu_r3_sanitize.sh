set +e
O=gpurun_out/r3
mkdir -p $O
K='test_attn2_tcgen05_against_the_oracle_core and fp16 or test_attn2_attn_drop_masks_match_the_oracle or test_attention_backward_tcgen05'
timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_attn2.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -x -q -k "$K" > $O/memcheck_attn2.out 2>&1
echo "memcheck rc=$?" >> $O/memcheck_attn2.out
tail -3 $O/memcheck_attn2.out; tail -4 $O/memcheck_attn2.log
timeout 1200 compute-sanitizer --tool racecheck --log-file $O/racecheck_attn2.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -x -q -k "$K" > $O/racecheck_attn2.out 2>&1
echo "racecheck rc=$?" >> $O/racecheck_attn2.out
tail -3 $O/racecheck_attn2.out; tail -6 $O/racecheck_attn2.log
