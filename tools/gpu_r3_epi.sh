set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_epi.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_epi.log
tail -4 $O/gputests_epi.log
for v in 256 128; do
DPMN_TC_BN_MAX=$v timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_epi_bn$v.json 2> $O/bench_epi_bn$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_epi_bn{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('BN_MAX',sys.argv[1],d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['by_kernel_ms'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"gemm_tc_kernel|gemm_res_ln|attn2_tc|mlp_fc1_dw|sk_gate_c96" --launch-skip 160 -c 16 --csv --log-file $O/block_times_epi.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-psn > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r3/block_times_epi.csv') if l.startswith('"')))
h=rows[0]; i_n=h.index('Kernel Name'); i_m=h.index('Metric Name'); i_v=h.index('Metric Value'); i_id=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[i_id], r[i_n][:45]),{})[r[i_m]]=r[i_v]
for k,v in d.items(): print(k, v)
PY
