# usage: gpu_r2_ncu.sh <kernel-regex> <out-name> [count] [skip]
set +e
mkdir -p gpurun_out/r2
K="$1"; O="$2"; C="${3:-1}"; S="${4:-0}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $S -c $C -o gpurun_out/r2/$O -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2/$O.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/r2/$O.log; ls -la gpurun_out/r2/$O.ncu-rep
