set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/multi_probe.py > gpurun_out/r2_multi_probe_n$NG.log 2> gpurun_out/r2q_multi_probe.err; echo "rc=$?"
cat gpurun_out/r2_multi_probe_n$NG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 200 --warmup 3 --no-dropin > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r2q_bench.err; echo "rc=$?"
tail -n 3 gpurun_out/r2q_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$NG.json'))
print({k:d[k] for k in ('n_gpus','ms_per_step','value','fill_only_ms_per_step','nccl_allgather_ms_per_step','gather_check','kernel_ms_per_launch','clocks')})
print(d['e2e']); print(d['acim'])
PY
if [ "$NG" = "8" ]; then
timeout 300 python -m pytest tests/test_gpu_group.py "tests/test_gpu_transfer.py::test_plain_c_group_client" -q -s 2>&1 | tail -n 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 tools/sweep_bench.py > gpurun_out/r2_sweep_config5_n$NG.log 2>/dev/null; tail -n 1 gpurun_out/r2_sweep_config5_n$NG.log
fi
