set +e
O=gpurun_out/r3
mkdir -p $O
timeout 600 python -m pytest tests/test_psn.py -x -q > $O/psn_test.log 2>&1
echo "rc=$?" >> $O/psn_test.log
tail -6 $O/psn_test.log
timeout 600 python bench.py --no-cpu-baseline --no-train > $O/bench_psn.json 2> $O/bench_psn.err
tail -3 $O/bench_psn.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_psn.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d.get('psn'), indent=1))
PY
