#!/bin/bash
# bench variants of the C2 step with different sets of sequence-resident kernels (ADT_SEQ_FUSED bit mask) -> gpurun_out/variants.txt
for m in 0 1 17 31; do
  ADT_SEQ_FUSED=$m python bench.py --steps 20 --warmup 5 --skip c1,c5,refgpu,fp32,selfcheck,evo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mask', $m, 'ms/step %.4f' % d['ms_per_step'], 'seqs/s %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'launches', d['gpu_launches_per_step'], 'eval %.0f' % d['eval']['value'], 'eval_e2e %.0f' % d['eval']['e2e']['value'])
"
done
