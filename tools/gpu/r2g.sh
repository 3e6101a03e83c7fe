set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_group.py -x -q -s > gpurun_out/r2g_group.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_group.log
tail -n 15 gpurun_out/r2g_group.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "rc=$?"
tail -n 20 gpurun_out/r2g_bench_n2.err
cat gpurun_out/r2g_bench_n2.json
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "rc=$?"
tail -n 5 gpurun_out/r2g_bench_n1.err
cat gpurun_out/r2g_bench_n1.json
