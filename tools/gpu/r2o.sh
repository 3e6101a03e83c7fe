set -x
mkdir -p gpurun_out
timeout 300 python tools/api_probe.py > gpurun_out/r2o_api.log 2>&1
cat gpurun_out/r2o_api.log | tr '}' '\n' | cut -c1-200
