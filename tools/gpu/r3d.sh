set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -n 3 $O/r2_gpu_tests.log
timeout 600 python bench.py --steps 200 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_ref.err; echo "ref rc=$?"
timeout 120 python tools/ncu_fill.py > $O/r2_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_launches.csv python tools/ncu_fill.py > $O/r2_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"basis_mma|transform_kernel|invert_branches" -s 6 -c 3 -f -o $O/r2_final python tools/ncu_fill.py > $O/r2_ncu2.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin --no-graph > $O/r2_plainb.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_py.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin --no-graph > $O/r2_ncu3.log 2>&1
QT_CHOP=1 QT_SPARSE=1 timeout 120 python tools/quick_time.py > $O/r2_quick_time_configs.log 2>&1
timeout 120 python tools/api_probe.py > $O/r2_api_latency.log 2>&1
timeout 120 python tools/dropin_probe.py > $O/r2_dropin_probe.log 2>&1
timeout 300 python tools/sweep_bench.py > $O/r2_sweep_config5_n1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
cat $O/r2_bench_n1.json | cut -c1-700
