#!/bin/bash
# Dev probe: ncu source-level capture of the pipelined linear backward (LayerNorm variant, then the plain one).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lin_tc_bwd2 -c 1 -f -o gpurun_out/bwd2_ln python tests/probe/lin_bench.py > gpurun_out/bwd2_ncu_ln.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lin_tc_bwd2 --launch-skip 13 -c 1 -f -o gpurun_out/bwd2_plain python tests/probe/lin_bench.py > gpurun_out/bwd2_ncu_plain.log 2>&1
ls -la gpurun_out/*.ncu-rep
