set -x
mkdir -p gpurun_out
timeout 600 python tools/adaptive_agreement.py > gpurun_out/r2j_adaptive.log 2>&1
cat gpurun_out/r2j_adaptive.log | cut -c1-330
