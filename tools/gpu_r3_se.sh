set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_se.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_se.log
tail -5 $O/gputests_se.log
timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_se.json 2> $O/bench_train_se.err
tail -3 $O/bench_train_se.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_train_se.json').read().strip().splitlines()[-1])
print(d.get('ms_per_step'), d.get('value'), d['roofline'].get('by_kernel_ms') if 'roofline' in d else None)
PY
