set +e
O=gpurun_out/r3; mkdir -p $O
for v in 1 0 1 0; do
DPMN_REDUCE_TILED=$v timeout 400 python bench.py --mode train --no-cpu-baseline --steps 10 --warmup 3 > $O/bench_train_red$v.json 2> $O/bench_train_red$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_train_red{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('tiled',sys.argv[1],d.get('ms_per_step'), d['roofline']['by_kernel_ms'].get('bwd_conv'))
PY
done
