set +e
O=gpurun_out/r3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graphed_hot_path" > $O/gputests_graph32.log 2>&1
echo "rc=$?" >> $O/gputests_graph32.log
tail -15 $O/gputests_graph32.log
