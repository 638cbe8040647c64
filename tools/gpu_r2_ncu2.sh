set +e
mkdir -p gpurun_out/r2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|gemm_res_ln|attn_tc|mlp_fc1_dw|sk_gate_c96" --launch-skip 160 -c 16 -o gpurun_out/r2/ncu_block_v3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2/ncu_block_v3.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2/ncu_block_v3.ncu-rep
for s in 2 3; do
DPMN_BENCH_SLOTS=$s timeout 300 python bench.py --no-cpu-baseline --no-train > gpurun_out/r2/bench_slots$s.json 2> gpurun_out/r2/bench_slots$s.err
python - $s <<'PY'
import json, sys
try:
    d=json.loads(open(f"gpurun_out/r2/bench_slots{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("slots", sys.argv[1], d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("ERR", e)
PY
done
