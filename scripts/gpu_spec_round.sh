#!/bin/bash
# one gpurun call: the memo-kernel build variants of scripts/build_variants.py on the headline workload, one process each
# (a variant that hangs costs its own time limit only); the speculative-key variants are skipped once one of them fails
mkdir -p gpurun_out
L=gpurun_out/r2_spec_experiment.log
: > $L
run() { timeout $1 python scripts/spec_experiment.py 10000000 $2 >> $L 2>> gpurun_out/r2_spec_experiment.err; rc=$?; echo "{\"lib\": \"$2\", \"rc\": $rc}" >> $L; return $rc; }
run 75 base
run 35 k2
if run 40 spec2; then
  run 35 spec1
  run 35 spec2b
  run 35 k2spec2
fi
cat $L
