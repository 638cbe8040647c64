set +e
O=gpurun_out/r3
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "windowed or attn2" > $O/attn2_test_d.log 2>&1
echo "attn2 tests rc=$?" >> $O/attn2_test_d.log
tail -12 $O/attn2_test_d.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_d.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_d.log
tail -4 $O/gputests_d.log
timeout 600 python tools/attn_sweep.py > $O/sweep_full_v3.md 2>$O/sweep_full_v3.err
tail -30 $O/sweep_full_v3.md; tail -3 $O/sweep_full_v3.err
