#!/bin/bash
# Dev probe: A/B of two versions of lin_tc.cu on the same box (tests/probe/tmp/lin_tc_prev.cu.txt = version A).
cd "$GRAFT_REPO_ROOT" || exit 1
echo "--- B (current)"; timeout 300 python tests/probe/lin_bench.py 2>&1 | grep "bwd_ln\|bwd 32   \|bwd gelu\|bwd 96"
cp vaesne-dev_b200/csrc/lin_tc.cu /tmp/cur.cu; cp tests/probe/tmp/lin_tc_prev.cu.txt vaesne-dev_b200/csrc/lin_tc.cu
python vaesne-dev_b200/build.py --force > /dev/null 2>&1
echo "--- A (previous)"; timeout 300 python tests/probe/lin_bench.py 2>&1 | grep "bwd_ln\|bwd 32   \|bwd gelu\|bwd 96"
cp /tmp/cur.cu vaesne-dev_b200/csrc/lin_tc.cu; python vaesne-dev_b200/build.py --force > /dev/null 2>&1
echo "--- B again"; timeout 300 python tests/probe/lin_bench.py 2>&1 | grep "bwd_ln\|bwd 32   \|bwd gelu\|bwd 96"
