set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_n2_final.json 2> $O/bench_n2_final.err
echo rc=$?
tail -3 $O/bench_n2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_n2_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('psn',d['psn'] and d['psn']['value_psn_included'])
t=d['train']; print('train',t['ms_per_step'],t['value'],t['allreduce'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_n2_ref.json 2> $O/bench_n2_ref.err
echo rc=$?; cut -c1-300 $O/bench_n2_ref.json
