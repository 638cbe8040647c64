set +e
mkdir -p gpurun_out/r2
NG=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 50 --warmup 5 > gpurun_out/r2/bench_n${NG}.json 2> gpurun_out/r2/bench_n${NG}.err
echo "bench n$NG rc=$?"; tail -n 8 gpurun_out/r2/bench_n${NG}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 3 --no-overlap --train-steps 10 > gpurun_out/r2/bench_n${NG}_noov.json 2> gpurun_out/r2/bench_n${NG}_noov.err
python - $NG <<'PY'
import json, sys
ng=sys.argv[1]
for f in (f"bench_n{ng}", f"bench_n{ng}_noov"):
    try:
        d=json.loads(open(f"gpurun_out/r2/{f}.json").read().strip().splitlines()[-1])
        t=d["train"]; print(f, "infer", round(d["ms_per_step"],3), round(d["value"]), "| train", round(t["ms_per_step"],2), round(t["value"]), t["allreduce"])
    except Exception as e:
        print(f, "ERR", e)
PY
