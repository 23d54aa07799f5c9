#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none -k regex:layernorm -s 20 -c 3 -o gpurun_out/prof_ln_fwd \
   python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_a.log 2>&1
echo "== ncu ln fwd rc=$?"
timeout -k 10 600 ncu --set full --clock-control none -k regex:layernorm_bwd -s 40 -c 3 -o gpurun_out/prof_ln_bwd \
   python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_b.log 2>&1
echo "== ncu ln bwd rc=$?"
timeout -k 10 600 ncu --set full --clock-control none -k regex:groupnorm -s 66 -c 6 -o gpurun_out/prof_gn \
   python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_c.log 2>&1
echo "== ncu gn rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
