set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_fork.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_fork.log
tail -5 $O/gputests_fork.log
for v in 1 0; do
DPMN_CMM_FORK=$v timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_fork$v.json 2> $O/bench_train_fork$v.err
tail -2 $O/bench_train_fork$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_train_fork{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('fork',sys.argv[1],d.get('ms_per_step'), d.get('value'))
PY
done
