set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests_splitk.log 2>&1
echo "gpu tests rc=$?" >> $O/gputests_splitk.log
tail -4 $O/gputests_splitk.log
DPMN_CONV_SPLITK=0 timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_sk0.json 2> $O/bench_sk0.err
timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_sk1.json 2> $O/bench_sk1.err
python - <<'PY'
import json
for f in ('bench_sk0','bench_sk1'):
    d=json.loads(open(f'gpurun_out/r3/{f}.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['value'], 'conv_tc', d['roofline']['by_kernel_ms'].get('conv_tc'))
PY
tail -2 $O/bench_sk1.err
