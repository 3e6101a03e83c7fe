set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_group.py -x -q -s > gpurun_out/r2f_group.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_group.log
tail -n 30 gpurun_out/r2f_group.log
