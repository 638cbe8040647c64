set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_red.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_red.log
tail -5 $O/gputests_red.log
timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_red.json 2> $O/bench_train_red.err
tail -2 $O/bench_train_red.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_train_red.json').read().strip().splitlines()[-1])
print(d.get('ms_per_step'), d.get('value'), d['roofline'].get('by_kernel_ms') if 'roofline' in d else None)
PY
