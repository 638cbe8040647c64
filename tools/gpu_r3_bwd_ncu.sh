set +e
O=gpurun_out/r3
mkdir -p $O
python tools/attn_bwd_one.py 48 20 0.1
python tools/attn_bwd_one.py 48 20 0.0
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:attn2_bwd --launch-skip 3 --launch-count 1 -o $O/ncu_attn2_bwd_v1 -f python tools/attn_bwd_one.py 48 3 0.1 > $O/ncu_attn2_bwd_v1.log 2>&1
tail -2 $O/ncu_attn2_bwd_v1.log
