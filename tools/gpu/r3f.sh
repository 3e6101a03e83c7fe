mkdir -p gpurun_out
L=gpurun_out/r2_shard_probe.log
: > $L
timeout 200 python tools/shard_probe.py 2>&1 | grep "^{" >> $L
PGB_KBM_SC16=1 timeout 200 python tools/shard_probe.py 2>&1 | grep "^{" >> $L
cat $L
