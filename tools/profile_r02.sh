#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU): launch list of a short bench run + full captures of the dominant kernels.
# The .ncu-rep files are exported to CSV on the box and removed (gpurun brings back at most 64 MiB).
SK="--skip c1,c3,c5,refgpu,fp32,evo,selfcheck --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 $SK > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attn_small_bwd|post_bwd_small|pre_bwd_small2|mid_bwd_small|enc_seq_fwd" -s 60 -c 10 -o gpurun_out/r02_train -f python bench.py --steps 2 --warmup 3 $SK > gpurun_out/r02_train_ncu.log 2>&1
K7_ONLY=512,256 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel|rescore_select" -s 6 -c 3 -o gpurun_out/r02_k7 -f python tools/k7_bench.py > gpurun_out/r02_k7_ncu.log 2>&1
K7_ONLY=4096,256 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel" -s 2 -c 1 -o gpurun_out/r02_k7_4096 -f python tools/k7_bench.py >> gpurun_out/r02_k7_ncu.log 2>&1
for f in r02_train r02_k7 r02_k7_4096; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
  rm -f gpurun_out/$f.ncu-rep
done
ls -la gpurun_out/
