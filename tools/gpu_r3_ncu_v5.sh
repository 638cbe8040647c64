set +e
O=gpurun_out/r3; mkdir -p $O
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel|gemm_res_ln|attn2_tc|mlp_fc1_dw|sk_gate_c96" --launch-skip 160 -c 16 -o $O/ncu_block_v5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-psn > /dev/null 2>&1
ls -la $O/ncu_block_v5.ncu-rep
python tools/ncu_summary.py $O/ncu_block_v5.ncu-rep > $O/ncu_block_v5.txt
cat $O/ncu_block_v5.txt | cut -c1-170
