set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_final3.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_final3.log
tail -4 $O/gputests_final3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_final3.log 2>&1
tail -2 $O/smoke_final3.log
timeout 900 python bench.py > $O/bench_final_v4.json 2> $O/bench_final_v4.err
echo rc=$?
tail -2 $O/bench_final_v4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_final_v4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['parity']['max_abs_err_over_max_ref'])
print(d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['attention_roofline']['frac_of_attention_flop_roofline'], d['roofline']['by_kernel_ms'])
print(d['psn']['ms_per_step_psn_included'], d['psn']['value_psn_included'])
print(d['train']['ms_per_step'], d['train']['value'], d['train']['by_kernel_ms'])
PY
