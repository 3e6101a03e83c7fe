mkdir -p gpurun_out
L=gpurun_out/r2_e2e_ragged_tail_probe.log
: > $L
timeout 300 python -m pytest tests/test_gpu_transfer.py -x -q -k "grid_cap" 2>&1 | tail -n 3
for w in "" "256,256,512,512,1024,1024,2048" "512,512,512,512,1024,1024,1024,2048" "512,1024,1024,2048" "1024,1024,2048" "256,512,1024,2048"; do
  echo "# RAGGED_TAIL=1 PGB_TRACE_RAGGED=1 PGB_RAGGED_TAIL_WIDTHS=$w" >> $L
  if [ -z "$w" ]; then
    RAGGED_TAIL=1 PGB_TRACE_RAGGED=1 timeout 200 python tools/e2e_ragged_probe.py 2>&1 | tail -n 2 >> $L
    RAGGED_TAIL=1 timeout 200 python tools/e2e_ragged_probe.py 2>&1 | tail -n 1 >> $L
  else
    RAGGED_TAIL=1 PGB_TRACE_RAGGED=1 PGB_RAGGED_TAIL_WIDTHS=$w timeout 200 python tools/e2e_ragged_probe.py 2>&1 | tail -n 2 >> $L
    RAGGED_TAIL=1 PGB_RAGGED_TAIL_WIDTHS=$w timeout 200 python tools/e2e_ragged_probe.py 2>&1 | tail -n 1 >> $L
  fi
done
cat $L
