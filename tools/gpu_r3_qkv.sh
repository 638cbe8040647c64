set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_hint.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_hint.log
tail -4 $O/gputests_hint.log
timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_hint.json 2> $O/bench_hint.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_hint.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['by_kernel_ms'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"gemm_tc_kernel|gemm_res_ln|attn2_tc|mlp_fc1_dw|sk_gate_c96" --launch-skip 160 -c 16 --csv --log-file $O/block_times_hint.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-psn > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r3/block_times_hint.csv') if l.startswith('"')))
h=rows[0]; i_n=h.index('Kernel Name'); i_m=h.index('Metric Name'); i_v=h.index('Metric Value'); i_id=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[i_id], r[i_n][:45]),{})[r[i_m]]=r[i_v]
for k,v in d.items(): print(k, v)
PY
