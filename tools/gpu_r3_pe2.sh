set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_pe2.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_pe2.log
tail -4 $O/gputests_pe2.log
for v in 2 1 2 1; do
DPMN_PATCH_EMBED_TOKENS=$v timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_pe2_$v.json 2> $O/bench_pe2_$v.err
python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3/bench_pe2_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('tokens/lane-group',sys.argv[1],d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['by_kernel_ms']['patch_embed'])
PY
done
