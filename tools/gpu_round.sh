#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, launch list.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
for f in test_gemm_gpu test_ops_gpu test_models_gpu; do
  timeout -k 10 300 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/$f.full 2>&1; rc=$?
  cut -c1-400 gpurun_out/$f.full > gpurun_out/$f.log; rm -f gpurun_out/$f.full
  echo "== $f (rc=$rc): $(tail -1 gpurun_out/$f.log)"
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "== $f TIMED OUT: aborting the rest of the session to save GPU budget"; exit 1; fi
done
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout -k 10 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
     python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
  echo "== ncu rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
fi
if [ "$2" == "full" ]; then
  # --set full captures: 8 GEMMs of one mixer layer's backward, then 8 GEMMs of the CLIP forward
  timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 448 -c 8 -o gpurun_out/prof_gemm_mixer_bwd \
     python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof1.log 2>&1
  echo "== ncu full mixer-bwd rc=$?"
  timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 215 -c 8 -o gpurun_out/prof_gemm_clip \
     python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof2.log 2>&1
  echo "== ncu full clip rc=$?"
fi
