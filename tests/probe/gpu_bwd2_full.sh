#!/bin/bash
# Dev probe: whole GPU suite (default dispatch, then the pipelined linear backward forced on for every eligible call), then the default bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
VAESNE_LIN_BWD2=2 timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_bwd2.json 2> gpurun_out/bench_bwd2.err; tail -c 300 gpurun_out/bench_bwd2.json; tail -3 gpurun_out/bench_bwd2.err
