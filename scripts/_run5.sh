mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6_tests.log 2>&1; tail -15 gpurun_out/s6_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-other-mode --kernels-out gpurun_out/s6_kernels.json > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; tail -3 gpurun_out/s6_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-other-mode --dtype fp32 --steps 20 > gpurun_out/s6_bench_fp32.json 2>> gpurun_out/s6_bench.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s6_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d['gpu_launches'], d['kernel_families_ms_per_step'])
    except Exception as e: print(f,'ERR',e)
P
