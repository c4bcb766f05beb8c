#!/bin/bash
# one gpurun call: memo-kernel build variants (scripts/build_variants.py) on the headline workload, one process each (a
# variant that hangs costs its own time limit only), then the smoke test of the library as shipped
mkdir -p gpurun_out
L=gpurun_out/r2_spec_experiment2.log
: > $L
run() { timeout $1 python scripts/spec_experiment.py 10000000 $2 >> $L 2>> gpurun_out/r2_spec_experiment2.err; rc=$?; echo "{\"lib\": \"$2\", \"rc\": $rc}" >> $L; return $rc; }
run 40 final
run 35 spec2n
run 35 spec2b
cat $L
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
