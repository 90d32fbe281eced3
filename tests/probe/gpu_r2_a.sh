#!/bin/bash
# round-2 GPU session A: parity of the new range management / full-machine test, timing of the attention kernels, sanitizer logs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
ONLY_TIME=1 python tests/probe/attn_tc_check.py > gpurun_out/r2a_attn_time.log 2>&1
cat gpurun_out/r2a_attn_time.log
python tests/probe/lin_bench.py > gpurun_out/r2a_lin_bench.log 2>&1
tail -30 gpurun_out/r2a_lin_bench.log
