set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.log
grep -E "passed|failed|rc=|C ABI|equal|group of" gpurun_out/r2k_tests.log | tail -n 40
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NG --steps 100 --warmup 3 --no-dropin > gpurun_out/r2k_bench_n$NG.json 2> gpurun_out/r2k_bench.err; echo "rc=$?"
tail -n 3 gpurun_out/r2k_bench.err
python - <<'PY'
import json,glob
for f in glob.glob('gpurun_out/r2k_bench_n*.json'):
    d=json.load(open(f)); print(f, d['ms_per_step'], d['fill_only_ms_per_step'], d['e2e']['ms_per_step'], d['gather_check'], d['kernel_ms_per_launch'], d['clocks'])
PY
