set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_wstage.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_wstage.log
tail -4 $O/gputests_wstage.log
timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_wstage.json 2> $O/bench_train_wstage.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_train_wstage.json').read().strip().splitlines()[-1])
k=d['roofline']['by_kernel_ms']
print(d.get('ms_per_step'), 'gemm',k.get('gemm'),'layernorm',k.get('layernorm'),'convert',k.get('convert'))
PY
