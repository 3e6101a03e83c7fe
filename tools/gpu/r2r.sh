mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for f in 0 1 2 3; do
PGB_MC_FLAGS=$f timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2952$f tools/multi_probe_short.py 2>/dev/null | tee -a gpurun_out/r2_mc_flags_n$NG.log
done
