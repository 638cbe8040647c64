set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_pwfork.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_pwfork.log
tail -5 $O/gputests_pwfork.log
for v in 1 0; do
DPMN_PGRM_WGRAD_FORK=$v timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_pwfork$v.json 2> $O/bench_train_pwfork$v.err
tail -2 $O/bench_train_pwfork$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_train_pwfork{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('pgrm wgrad fork',sys.argv[1],d.get('ms_per_step'), d.get('value'))
PY
done
