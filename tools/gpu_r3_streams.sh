set +e
O=gpurun_out/r3
mkdir -p $O
run() { # name one_stream slots pdl
  DPMN_BENCH_ONE_STREAM=$2 DPMN_BENCH_SLOTS=$3 DPMN_PDL=$4 timeout 300 python bench.py --no-cpu-baseline --no-train --no-psn > $O/bench_st_$1.json 2> $O/bench_st_$1.err
  python - $1 <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/r3/bench_st_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1],round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run s3_k2_p0 0 2 0
run s1_k2_p0 1 2 0
run s1_k3_p0 1 3 0
run s1_k4_p0 1 4 0
run s1_k2_p1 1 2 1
run s1_k3_p1 1 3 1
run s1_k4_p1 1 4 1
run s3_k3_p0 0 3 0
run s3_k3_p1 0 3 1
