set +e
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -x -q -k "cmm" 2>&1 | tail -n 6
timeout 300 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2/bench_train_$1.json 2> gpurun_out/r2/bench_train_$1.err
python - $1 <<'PY'
import json, sys
try:
    d=json.loads(open(f"gpurun_out/r2/bench_train_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("train", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"]); print(d["train"]["by_kernel_ms"])
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/r2/bench_train_{sys.argv[1]}.err").read()[-2000:])
PY
