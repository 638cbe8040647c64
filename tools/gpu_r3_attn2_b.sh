set +e
O=gpurun_out/r3
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "windowed or attn2" > $O/attn2_test_b.log 2>&1
echo "attn2 tests rc=$?" >> $O/attn2_test_b.log
tail -4 $O/attn2_test_b.log
timeout 300 python tools/attn_sweep.py --quick > $O/sweep_v2h.md 2>$O/sweep_v2h.err
cat $O/sweep_v2h.md; tail -3 $O/sweep_v2h.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:attn2 --launch-skip 3 --launch-count 1 -o $O/ncu_attn2_v2 -f python tools/attn_one.py 48 6 > $O/ncu_attn2_v2.log 2>&1
tail -2 $O/ncu_attn2_v2.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_b.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_b.log
tail -4 $O/gputests_b.log
timeout 300 python bench.py --no-cpu-baseline --no-train > $O/bench_v2b.json 2> $O/bench_v2b.err
cut -c1-300 $O/bench_v2b.json
