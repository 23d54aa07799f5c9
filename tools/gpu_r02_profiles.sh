#!/bin/bash
# round 2 evidence session (one GPU): ncu launch list of one step (durations + DRAM bytes), ncu --set full of the HBM-bound kernels and
# the new attention kernels, the four bench lines, the reference arm, the eager-GPU context number
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_pytest_gpu_final.log)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/one_step.py > gpurun_out/r02_one_step.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/r02_launches.csv) lines"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"layernorm_fwd_pipe|layernorm_bwd_pipe|groupnorm_apply|groupnorm_bwd_apply|groupnorm_stats|adam_vec4|warp_bwd|warp_fwd|cutout_final|pool_bwd|pool_fwd|mha_small|colsum|rowsum|conv_taps_gather|vq_" \
    -c 60 -o gpurun_out/r02_ncu_hbm_kernels python tools/one_step.py > gpurun_out/r02_ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
ncu -i gpurun_out/r02_ncu_hbm_kernels.ncu-rep --page raw --csv > gpurun_out/r02_ncu_hbm_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_hbm_kernels.ncu-rep      # 60 kernels x --set full + source: far beyond the 64 MiB that travel back
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"mha_flash" -c 4 \
    -o gpurun_out/r02_ncu_flash python tools/one_step.py --config 4 > gpurun_out/r02_ncu_flash.log 2>&1
echo "ncu flash rc=$?"
ncu -i gpurun_out/r02_ncu_flash.ncu-rep --page raw --csv > gpurun_out/r02_ncu_flash_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
for c in 2 3 4 5; do
  extra=""; [ $c = 2 ] && extra="--gpu-eager-baseline"
  timeout 600 python bench.py --config $c --steps 20 --warmup 4 $extra > gpurun_out/r02_bench_final_c$c.json 2> gpurun_out/r02_bench_final_c$c.err
  echo "bench config $c rc=$? $(tail -c 300 gpurun_out/r02_bench_final_c$c.err | tr '\n' ' ')"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_c2.json 2>/dev/null; echo "reference rc=$?"
