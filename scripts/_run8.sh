set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -x -q > gpurun_out/s8_nccl_tests.log 2>&1; tail -3 gpurun_out/s8_nccl_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-other-mode > gpurun_out/s8_n2.json 2> gpurun_out/s8_n2.err
timeout 400 $TR bench.py --gpus 2 --steps 50 --warmup 5 --impl reference > gpurun_out/s8_n2_ref.json 2> gpurun_out/s8_n2_ref.err || true
python - <<'P'
import json
d=json.loads(open('gpurun_out/s8_n2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['comm_sms'])
P
tail -c 600 gpurun_out/s8_n2_ref.json
