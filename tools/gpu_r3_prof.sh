set +e
O=gpurun_out/r3
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|gemm_res_ln|attn2_tc|mlp_fc1_dw|sk_gate_c96" --launch-skip 160 -c 16 -o $O/ncu_block_v4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-psn > $O/ncu_block_v4.log 2>&1
echo "ncu rc=$?"; ls -la $O/ncu_block_v4.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-psn > $O/ll_v3.log 2>&1
echo rc=$?
python tools/launch_list.py $O/launches_v3.csv | head -n 30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_train_v3.csv python tools/one_train_step.py > $O/llt_v3.log 2>&1
echo rc=$?
python tools/launch_list.py $O/launches_train_v3.csv | head -n 45
