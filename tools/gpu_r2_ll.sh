set +e
mkdir -p gpurun_out/r2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2/launches_$1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2/ll_$1.log 2>&1
echo rc=$?
python tools/launch_list.py gpurun_out/r2/launches_$1.csv | head -n 40
