set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_bn.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_bn.log
tail -5 $O/gputests_bn.log
timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_bn.json 2> $O/bench_train_bn.err
tail -2 $O/bench_train_bn.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_train_bn.json').read().strip().splitlines()[-1])
print(d.get('ms_per_step'), d.get('value'), d['roofline'].get('by_kernel_ms') if 'roofline' in d else None)
PY
DPMN_TRAIN_STREAMS=0 DPMN_CMM_FORK=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 -c 2400 --csv --log-file $O/launches_train_v4.csv python bench.py --mode train --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
python tools/launch_list.py $O/launches_train_v4.csv > $O/launches_train_v4.txt 2>&1
head -50 $O/launches_train_v4.txt
