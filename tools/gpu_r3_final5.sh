set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python bench.py > $O/bench_final_v5.json 2> $O/bench_final_v5.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_final_v5.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['parity']['max_abs_err_over_max_ref'])
print(d['roofline']['frac'], d['psn']['value_psn_included'], d['train']['ms_per_step'], d['train']['value'], d['train']['gpu_launches'])
PY
