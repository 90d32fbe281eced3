#!/bin/bash
# Dev probe: pipelined linear backward (lin_tc_bwd2) - parity with it forced on for every eligible shape, then timing.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
VAESNE_LIN_BWD2=2 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "lin" 2>&1 | tail -15
echo "--- old"; VAESNE_LIN_BWD2=0 timeout 300 python tests/probe/lin_bench.py 2>&1 | grep bwd
echo "--- new"; VAESNE_LIN_BWD2=1 timeout 300 python tests/probe/lin_bench.py 2>&1 | grep bwd
