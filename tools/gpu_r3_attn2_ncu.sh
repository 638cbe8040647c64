set +e
O=gpurun_out/r3
mkdir -p $O
timeout 300 python tools/attn_sweep.py --quick > $O/sweep_v2g.md 2>$O/sweep_v2g.err
DPMN_ATTN_V1=1 timeout 300 python tools/attn_sweep.py --quick > $O/sweep_v1g.md 2>$O/sweep_v1g.err
cat $O/sweep_v1g.md $O/sweep_v2g.md; tail -3 $O/sweep_v2g.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:attn2 --launch-skip 3 --launch-count 1 -o $O/ncu_attn2_v1 -f python tools/attn_one.py 48 6 > $O/ncu_attn2_v1.log 2>&1
tail -3 $O/ncu_attn2_v1.log
ls -la $O/*.ncu-rep
