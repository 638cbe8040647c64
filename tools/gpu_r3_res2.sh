set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_res2.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_res2.log
tail -4 $O/gputests_res2.log
timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_res2.json 2> $O/bench_res2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_res2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['by_kernel_ms'], d['parity']['max_abs_err_over_max_ref'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"gemm_res_ln" --launch-skip 48 -c 4 --csv --log-file $O/block_times_res2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-psn > /dev/null 2>&1
grep -o '"gemm_res_ln[^"]*".*' $O/block_times_res2.csv | awk -F'","' '{print $(NF-2), $NF}' | head -8
