set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_wfork.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_wfork.log
tail -5 $O/gputests_wfork.log
timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_wfork.json 2> $O/bench_train_wfork.err
tail -2 $O/bench_train_wfork.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_train_wfork.json').read().strip().splitlines()[-1])
print(d.get('ms_per_step'), d.get('value'))
PY
