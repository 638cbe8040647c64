set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_pdl.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_pdl.log
tail -4 $O/gputests_pdl.log
DPMN_PDL=0 timeout 300 python bench.py --no-cpu-baseline --no-train > $O/bench_pdl0.json 2> $O/bench_pdl0.err
timeout 300 python bench.py --no-cpu-baseline --no-train > $O/bench_pdl1.json 2> $O/bench_pdl1.err
cut -c1-260 $O/bench_pdl0.json; cut -c1-260 $O/bench_pdl1.json; tail -2 $O/bench_pdl1.err
timeout 300 python tools/attn_sweep.py --quick > $O/sweep_pdl.md 2>$O/sweep_pdl.err
tail -8 $O/sweep_pdl.md
