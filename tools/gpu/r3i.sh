mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/ncu_fill.py > $O/r3i_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"transform_kernel" -s 2 -c 1 -f -o $O/r2_kc_v6 python tools/ncu_fill.py > $O/r3i_ncu.log 2>&1
echo "ncu rc=$?"
