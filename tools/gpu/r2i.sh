set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_group.py -x -q -s > gpurun_out/r2i_group.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_group.log
tail -n 12 gpurun_out/r2i_group.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/multi_probe.py > gpurun_out/r2i_multi_probe_n$NG.log 2> gpurun_out/r2i_multi_probe.err; echo "rc=$?"
tail -n 3 gpurun_out/r2i_multi_probe.err; cat gpurun_out/r2i_multi_probe_n$NG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 100 --warmup 3 > gpurun_out/r2i_bench_n$NG.json 2> gpurun_out/r2i_bench.err; echo "rc=$?"
tail -n 5 gpurun_out/r2i_bench.err
cat gpurun_out/r2i_bench_n$NG.json
