#!/bin/bash
# Dev probe: phase timers (prologue / tile loop / flush) of lin_tc_bwd2_kernel; rebuilds with -DVAESNE_B2_PROF on the box.
cd "$GRAFT_REPO_ROOT" || exit 1
VAESNE_LIN_BWD2=2 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "lin" 2>&1 | tail -2
VAESNE_LIN_BWD2=0 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "lin" 2>&1 | tail -2
echo "--- old"; VAESNE_LIN_BWD2=0 timeout 300 python tests/probe/lin_bench.py 2>&1 | grep bwd
echo "--- new"; VAESNE_LIN_BWD2=1 timeout 300 python tests/probe/lin_bench.py 2>&1 | grep bwd
echo "--- new T=131072"; T=131072 VAESNE_LIN_BWD2=2 timeout 300 python tests/probe/lin_bench.py 2>&1 | grep bwd
VAESNE_B2_PROF=1 python vaesne-dev_b200/build.py --force > /dev/null 2>&1
for T in 16384 1005568; do echo "== T=$T"; T=$T VAESNE_LIN_BWD2=2 timeout 300 python tests/probe/lin_bench.py 2>&1 | grep "b2prof" | sed -n '30,35p'; done
