set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; tail -2 gpurun_out/f_tests.log
timeout 600 python bench.py --kernels-out gpurun_out/f_kernels_bf16.json > gpurun_out/f_bench_bf16.json 2> gpurun_out/f_bench_bf16.err
timeout 600 python bench.py --dtype fp32 --steps 20 --no-cpu-baseline > gpurun_out/f_bench_fp32.json 2> gpurun_out/f_bench_fp32.err
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_bf16.csv python bench.py --steps 3 --warmup 3 --no-graph --no-other-mode --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
timeout 600 python scripts/bench_stress.py gpurun_out/f_stress.json > gpurun_out/f_stress.log 2>&1
python -c "from __graft_entry__ import smoke; smoke(); print('smoke ok')" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
