set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -x -q > gpurun_out/s5_nccl_tests.log 2>&1; tail -3 gpurun_out/s5_nccl_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-other-mode > gpurun_out/s5_n2.json 2> gpurun_out/s5_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --batch 1024 --no-cpu-baseline --no-other-mode > gpurun_out/s5_n2_b1024.json 2> gpurun_out/s5_n2_b1024.err
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 --comm-sms 8 --no-cpu-baseline --no-other-mode > gpurun_out/s5_n2_c8.json 2> gpurun_out/s5_n2_c8.err
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 --comm-sms 0 --no-cpu-baseline --no-other-mode > gpurun_out/s5_n2_c0.json 2> gpurun_out/s5_n2_c0.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s5_n2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
P
