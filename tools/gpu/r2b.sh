set -x
mkdir -p gpurun_out
./tools/fp64_latency_probe > gpurun_out/r2b_fp64_latency.json 2>&1
python tools/ncu_fill.py > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:basis_mma -s 1 -c 1 -o gpurun_out/r2b_kbm python tools/ncu_fill.py > gpurun_out/r2b_ncu.log 2>&1
cat gpurun_out/r2b_fp64_latency.json
