#!/bin/bash
# Dev probe: whole GPU suite with the pipelined linear backward forced on for every eligible call, then the default bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
VAESNE_LIN_BWD2=2 timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/bench_bwd2.json 2> gpurun_out/bench_bwd2.err; tail -c 2500 gpurun_out/bench_bwd2.json
