set +e
O=gpurun_out/r3
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_train_tail.py tests/test_gpu_backward.py -m gpu -x -q > $O/gputests_t2s.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_t2s.log
tail -15 $O/gputests_t2s.log
for v in 1 0; do
DPMN_TRAIN_STREAMS=$v timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_s$v.json 2> $O/bench_train_s$v.err
tail -3 $O/bench_train_s$v.err
python - $v <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/r3/bench_train_s{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print('streams',sys.argv[1],d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'))
except Exception as e: print('ERR',e)
PY
done
