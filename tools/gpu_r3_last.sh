set +e
O=gpurun_out/r3; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/gputests_last.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_last.log
tail -3 $O/gputests_last.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
