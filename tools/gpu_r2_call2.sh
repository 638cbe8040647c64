set +e
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gputests_c2.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2/gputests_c2.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2/bench_fusedA.json 2> gpurun_out/r2/bench_fusedA.err
DPMN_FUSED_MLP_A=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2/bench_nofusedA.json 2> gpurun_out/r2/bench_nofusedA.err
tail -n 5 gpurun_out/r2/gputests_c2.log
python - <<'PY'
import json
for f in ("bench_fusedA","bench_nofusedA"):
    try:
        d=json.loads(open(f"gpurun_out/r2/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["by_kernel_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
