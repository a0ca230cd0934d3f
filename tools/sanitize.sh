#!/bin/bash
# compute-sanitizer summaries for profiles/ (run under gpurun): memcheck + racecheck of three training steps and an evaluation pass
# through (a) the default kernels, (b) every sequence-resident kernel, and memcheck of the catalog scorer / samplers.
out=gpurun_out/r02_sanitizer.txt
: > $out
run() { # name, tool, env..., cmd
  name=$1; tool=$2; shift; shift
  echo "== $name ($tool)" >> $out
  timeout 900 env "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|========= [A-Z]" | head -20 >> $out
}
run "train steps + eval, default kernels (bf16)" memcheck ADT_SEQ_FUSED=1 compute-sanitizer --tool memcheck --print-limit 5 python tools/seq_ab.py 2 50
run "train steps + eval, sequence-resident kernels" memcheck ADT_SEQ_FUSED=31 compute-sanitizer --tool memcheck --print-limit 5 python tools/seq_ab.py 2 50
run "train steps + eval, default kernels (bf16)" racecheck ADT_SEQ_FUSED=1 compute-sanitizer --tool racecheck --print-limit 5 python tools/seq_ab.py 2 50
run "train steps + eval, sequence-resident kernels" racecheck ADT_SEQ_FUSED=31 compute-sanitizer --tool racecheck --print-limit 5 python tools/seq_ab.py 2 50
run "catalog scorer 100k x 64 (two-pass tcgen05 + re-score), samplers" memcheck X=1 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_misc.py
cat $out
