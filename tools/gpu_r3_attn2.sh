set +e
mkdir -p gpurun_out/r3
O=gpurun_out/r3
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "windowed or attn2" > $O/attn2_test.log 2>&1
echo "attn2 tests rc=$?" >> $O/attn2_test.log
tail -15 $O/attn2_test.log
DPMN_ATTN_V1=1 timeout 300 python tools/attn_sweep.py --quick > $O/sweep_v1.md 2>$O/sweep_v1.err
timeout 300 python tools/attn_sweep.py --quick > $O/sweep_v2.md 2>$O/sweep_v2.err
cat $O/sweep_v1.md $O/sweep_v2.md
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests.log
tail -5 $O/gputests.log
DPMN_ATTN_V1=1 timeout 300 python bench.py --no-cpu-baseline --no-train > $O/bench_v1.json 2> $O/bench_v1.err
timeout 300 python bench.py --no-cpu-baseline --no-train > $O/bench_v2.json 2> $O/bench_v2.err
cut -c1-300 $O/bench_v1.json; cut -c1-300 $O/bench_v2.json
