set +e
mkdir -p gpurun_out/r2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/launches_train_$1.csv python tools/one_train_step.py > gpurun_out/r2/llt_$1.log 2>&1
echo rc=$?
python tools/launch_list.py gpurun_out/r2/launches_train_$1.csv | head -n 60
