set -x
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --warmup 5 --no-cpu-baseline --no-other-mode "$@" > gpurun_out/s5_$tag.json 2> gpurun_out/s5_$tag.err; }
run 8 n8_c16 --steps 50
run 8 n8_c0 --steps 50 --comm-sms 0
run 8 n8_c32 --steps 50 --comm-sms 32
run 4 n4_b512 --steps 30 --batch 512
run 4 n4_c0 --steps 50 --comm-sms 0
run 4 n4_c16 --steps 50
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s5_n[48]*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
P
