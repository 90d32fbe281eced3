#!/bin/bash
# Dev probe: accumulate-mode soak of the pipelined linear backward (1.0 M and 8.0 M tokens, all accumulate combinations, six rounds) with the in-kernel stage check (-DB2_CHECK build).
cd "$GRAFT_REPO_ROOT" || exit 1
B2_CHECK=1 python vaesne-dev_b200/build.py --force > /dev/null 2>&1
for i in 1 2 3 4 5 6; do MARK=1 TS=1005568,8044544 VAESNE_LIN_BWD2=2 timeout 300 python tests/probe/bwd2_acc_probe.py 2>&1 | grep "b2check\|^T=\|Error" | grep -v "e-07\|0.00e+00\|e-08" | cut -c1-200 | head -14; done
echo soak done
