set +e
O=gpurun_out/r3
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "windowed or attn2" > $O/attn2_test_c.log 2>&1
echo "attn2 tests rc=$?" >> $O/attn2_test_c.log
tail -4 $O/attn2_test_c.log
timeout 300 python tools/attn_sweep.py --quick > $O/sweep_v2i.md 2>$O/sweep_v2i.err
cat $O/sweep_v2i.md; tail -3 $O/sweep_v2i.err
timeout 300 python bench.py --no-cpu-baseline --no-train > $O/bench_v2c.json 2> $O/bench_v2c.err
cut -c1-300 $O/bench_v2c.json
