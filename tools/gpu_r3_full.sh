set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_full.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_full.log
tail -4 $O/gputests_full.log
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err
tail -3 $O/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_full.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'],'parity',d['parity'])
print('roofline',d['roofline']['kernel'],d['roofline']['frac'],'attn',d['roofline']['attention'])
print('psn',d['psn'] and (d['psn']['psn_ms_per_batch'],d['psn']['value_psn_included']))
print('train',d['train'] and (d['train']['ms_per_step'],d['train']['value']))
print(json.dumps(d['roofline']['by_kernel_ms']))
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
