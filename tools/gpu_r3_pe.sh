set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_pe.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_pe.log
tail -4 $O/gputests_pe.log
for v in new old; do
if [ $v = old ]; then export DPMN_PATCH_EMBED=4 DPMN_HEAD_MIX=1; fi
timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_pe_$v.json 2> $O/bench_pe_$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_pe_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1],d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['by_kernel_ms'], d['parity'])
PY
done
