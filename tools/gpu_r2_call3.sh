set +e
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gputests_c3.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2/gputests_c3.log
tail -n 25 gpurun_out/r2/gputests_c3.log
timeout 600 python bench.py --steps 50 > gpurun_out/r2/bench_c3.json 2> gpurun_out/r2/bench_c3.err
echo "bench rc=$?"; tail -n 5 gpurun_out/r2/bench_c3.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_c3.json").read().strip().splitlines()[-1])
    print("infer", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"])
    print("parity", d["parity"]); print("cpu", d["cpu_baseline"])
    r=d["roofline"]; print("roof", r["kernel"], r["frac"], r["pgrm_blocks"], r["attention"])
    t=d["train"]; print("train", t["ms_per_step"], t["value"], t["allreduce"], t.get("by_kernel_ms"))
except Exception as e:
    print("ERR", e)
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
