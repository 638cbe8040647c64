set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_head.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_head.log
tail -5 $O/gputests_head.log
for v in new old new; do
if [ $v = old ]; then export DPMN_HEAD_WGRAD_GROUPS=0 DPMN_HEAD_CONV1=1; else unset DPMN_HEAD_WGRAD_GROUPS DPMN_HEAD_CONV1; fi
timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_head_$v.json 2> $O/bench_train_head_$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_train_head_{sys.argv[1]}.json').read().strip().splitlines()[-1])
k=d['roofline']['by_kernel_ms']
print(sys.argv[1],d.get('ms_per_step'), 'bwd_head',k.get('bwd_head'),'head',k.get('head'))
PY
done
