mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for v in "PGB_MC_FLAGS=2" "PGB_MC_FLAGS=3" "PGB_SHARD_KA=1"; do
env $v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/multi_probe_short.py 2>/dev/null | sed "s/^{/{\"env\": \"$v\", /" | tee -a gpurun_out/r2_variants_n$NG.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 200 --warmup 3 --no-dropin > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r2t_bench.err; echo "rc=$?"
tail -n 2 gpurun_out/r2t_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$NG.json'))
print({k:d[k] for k in ('n_gpus','ms_per_step','value','fill_only_ms_per_step','nccl_allgather_ms_per_step','gather_check','kernel_ms_per_launch','clocks')})
print(d['e2e']['ms_per_step'], d['acim']['ms'])
PY
