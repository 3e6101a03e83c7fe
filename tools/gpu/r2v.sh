mkdir -p gpurun_out
L=gpurun_out/r2_e2e_ragged_regions.log
: > $L
timeout 600 python -m pytest tests/test_gpu_transfer.py tests/test_gpu_config4.py tests/test_gpu_parity.py -x -q -k "ragged or tail" 2>&1 | tail -n 3
run() { # regions widths
  echo "# RAGGED_TAIL=1 PGB_RAGGED_TAIL_REGIONS=$1 PGB_RAGGED_TAIL_WIDTHS=$2" >> $L
  RAGGED_TAIL=1 PGB_TRACE_RAGGED=1 PGB_RAGGED_TAIL_REGIONS=$1 PGB_RAGGED_TAIL_WIDTHS=$2 timeout 200 python tools/e2e_ragged_probe.py 2>&1 | grep "ragged trace" | tail -n 1 >> $L
  RAGGED_TAIL=1 PGB_RAGGED_TAIL_REGIONS=$1 PGB_RAGGED_TAIL_WIDTHS=$2 timeout 200 python tools/e2e_ragged_probe.py 2>&1 | tail -n 1 >> $L
}
run 1024,2048,2048,3072 512,512,1024,1024,1024,2048
run 512,1536,2048,4096 512,512,1024,1024,1024,2048
run 2048,2048,4096 512,512,1024,1024,1024,2048
run 1024,3072,4096 512,512,1024,1024,1024,2048
run 8192 512,512,1024,1024,1024,2048
run 1024,2048,2048,3072 512,512,1024,1024
run 1024,2048,2048,3072 256,256,512,1024,1024,1024,2048
run 768,1280,2048,4096 256,512,1280,1024
cat $L
