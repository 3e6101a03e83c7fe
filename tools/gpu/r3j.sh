mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 200 --warmup 3 --no-dropin > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r3j_bench.err; echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n$NG.json') if l.startswith('{')][0])
print({k:d[k] for k in ('n_gpus','ms_per_step','value','fill_only_ms_per_step','nccl_allgather_ms_per_step','gather_check','kernel_ms_per_launch')})
print(d['e2e']['ms_per_step'], d['acim']['ms'])
PY
