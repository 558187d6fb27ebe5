mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; python -c "import os;print('affinity',len(os.sched_getaffinity(0)),sorted(os.sched_getaffinity(0))[:8],'...')"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done; } > gpurun_out/s5_topo.txt 2>&1
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-other-mode > gpurun_out/s5_bind$i.json 2> gpurun_out/s5_bind$i.err
timeout 300 python bench.py --no-cpu-baseline --no-other-mode --no-numa-bind > gpurun_out/s5_nobind$i.json 2> gpurun_out/s5_nobind$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/s5_*bind*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d['config']['host_bound_to_gpu_numa_node'])
    except Exception as e: print(f,'ERR',e)
P
cat gpurun_out/s5_topo.txt
