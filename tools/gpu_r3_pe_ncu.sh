set +e
O=gpurun_out/r3
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"patch_embed|head_conv2" --launch-skip 24 -c 6 -o $O/ncu_pe_new -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-psn > /dev/null 2>&1
DPMN_PATCH_EMBED=4 DPMN_HEAD_MIX=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"patch_embed|head_conv2" --launch-skip 24 -c 6 -o $O/ncu_pe_old -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-psn > /dev/null 2>&1
ls -la $O/ncu_pe_*.ncu-rep
