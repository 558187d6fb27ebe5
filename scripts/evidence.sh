#!/bin/bash
# One GPU call that refreshes the measured evidence under gpurun_out/ (copied into profiles/ afterwards):
# ncu full captures of the hot kernels, launch list of a step, layer sweeps, bone-guided / stress / pool benches.
set -x
mkdir -p gpurun_out
LAYERS="0,32,16" REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"slab_gconv|slab_wgrad_kernel" -s 3 -c 3 -o gpurun_out/ev_l0 -f python scripts/bench_slab_layer.py > gpurun_out/ev_l0.log 2>&1
LAYERS="3,128,64" REPS=2 timeout 600 ncu --set full --clock-control none -k regex:"slab_conv_kernel|slab_wgrad_kernel" -s 3 -c 3 -o gpurun_out/ev_l3 -f python scripts/bench_slab_layer.py > gpurun_out/ev_l3.log 2>&1
LEVELS=0 timeout 600 ncu --set full --clock-control none -k regex:slab_pool -s 4 -c 2 -o gpurun_out/ev_pool -f python scripts/bench_slab_pool.py > gpurun_out/ev_pool.log 2>&1
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev_launches_bf16.csv python bench.py --steps 3 --warmup 3 --no-graph --no-other-mode --no-cpu-baseline > gpurun_out/ev_ncu_bench.log 2>&1
timeout 900 python scripts/sweep_layers.py --dtype bf16 > gpurun_out/ev_sweep_bf16.md 2> gpurun_out/ev_sweep_bf16.err
timeout 900 python scripts/sweep_layers.py --dtype fp32 > gpurun_out/ev_sweep_fp32.md 2> gpurun_out/ev_sweep_fp32.err
timeout 600 python scripts/bench_multiz.py > gpurun_out/ev_multiz.log 2>&1
timeout 600 python scripts/bench_stress.py gpurun_out/ev_stress.json > gpurun_out/ev_stress.log 2>&1
timeout 300 python scripts/bench_slab_pool.py > gpurun_out/ev_pool_bench.log 2>&1
timeout 300 python scripts/bench_pair_loss.py > gpurun_out/ev_pair_loss.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gl_ --csv --log-file gpurun_out/ev_heads_launches.csv python scripts/profile_heads.py > /dev/null 2>&1
DT=fp32 timeout 300 python scripts/bench_slab_pool.py > gpurun_out/ev_pool_bench_fp32.log 2>&1
ls -la gpurun_out/ev_*
