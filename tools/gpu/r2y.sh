mkdir -p gpurun_out
O=gpurun_out
QT_CHOP=1 QT_SPARSE=1 timeout 120 python tools/quick_time.py 2>&1 | cut -c1-330
timeout 120 python tools/ncu_fill.py > $O/r2y_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"basis_mma" -s 2 -c 1 -f -o $O/r2_kbm_v5 python tools/ncu_fill.py > $O/r2y_ncu.log 2>&1
echo "ncu rc=$?"; ls -la $O/r2_kbm_v5.ncu-rep
