set +e
mkdir -p gpurun_out/r2
DPMN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental.py -x -q 2>&1 | tail -n 8
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or full_hot_path or graphed" 2>&1 | tail -n 8
timeout 300 python bench.py --no-cpu-baseline --no-train > gpurun_out/r2/bench_kA.json 2> gpurun_out/r2/bench_kA.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_kA.json").read().strip().splitlines()[-1])
    print("infer", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"]); print(d["roofline"]["by_kernel_ms"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2/bench_kA.err").read()[-2000:])
PY
