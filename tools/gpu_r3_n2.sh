set +e
O=gpurun_out/r3; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-psn > $O/bench_n2_final_v3.json 2> $O/bench_n2_final_v3.err
echo rc=$?
tail -3 $O/bench_n2_final_v3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_n2_final_v3.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])
t=d['train']; print(t['ms_per_step'], t['value'], t['allreduce'], t.get('streams'))
PY
