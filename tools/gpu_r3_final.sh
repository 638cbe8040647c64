set +e
O=gpurun_out/r3; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_final.log 2>&1
tail -2 $O/smoke_final.log
timeout 900 python bench.py > $O/bench_final_v2.json 2> $O/bench_final_v2.err
echo rc=$?
tail -2 $O/bench_final_v2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_final_v2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['parity'])
print(d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['by_kernel_ms'])
print(d['psn']['ms_per_step_psn_included'], d['psn']['value_psn_included'])
print(d['train']['ms_per_step'], d['train']['value'], d['train']['by_kernel_ms'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_v2.json 2> $O/bench_ref_v2.err
tail -1 $O/bench_ref_v2.json | cut -c1-400
