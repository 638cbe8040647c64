set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_wsplit.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_wsplit.log
tail -4 $O/gputests_wsplit.log
for v in 1 0 1 0; do
DPMN_HEAD_WGRAD_SPLIT=$v timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_wsplit$v.json 2> $O/bench_train_wsplit$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_train_wsplit{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('split',sys.argv[1],d.get('ms_per_step'),'bwd_head',d['roofline']['by_kernel_ms'].get('bwd_head'))
PY
done
