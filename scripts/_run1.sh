set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s5_tests.log 2>&1; tail -3 gpurun_out/s5_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-other-mode > gpurun_out/s5_pdl_on.json 2> gpurun_out/s5_pdl_on.err
timeout 300 python bench.py --no-cpu-baseline --no-other-mode --no-dependent-launch > gpurun_out/s5_pdl_off.json 2> gpurun_out/s5_pdl_off.err
timeout 300 python bench.py --no-cpu-baseline --no-other-mode > gpurun_out/s5_pdl_on2.json 2>> gpurun_out/s5_pdl_on.err
timeout 300 python bench.py --no-cpu-baseline --no-other-mode --no-dependent-launch > gpurun_out/s5_pdl_off2.json 2>> gpurun_out/s5_pdl_off.err
timeout 300 python bench.py --no-cpu-baseline --no-other-mode --dtype fp32 --steps 20 > gpurun_out/s5_pdl_on_fp32.json 2>> gpurun_out/s5_pdl_on.err
for b in 512 1024 2048; do timeout 300 python bench.py --no-cpu-baseline --no-other-mode --batch $b --steps 20 > gpurun_out/s5_b${b}.json 2> gpurun_out/s5_b${b}.err; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
P
