set +e
O=gpurun_out/r3; mkdir -p $O
timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_var_train.json 2> $O/bench_var_train.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_var_train.json').read().strip().splitlines()[-1])
print('mode train:', d.get('ms_per_step'))
PY
timeout 600 python bench.py --no-cpu-baseline > $O/bench_var_default.json 2> $O/bench_var_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_var_default.json').read().strip().splitlines()[-1])
print('default: infer', d['ms_per_step'], 'train', d['train']['ms_per_step'])
PY
