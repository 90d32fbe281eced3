#!/bin/bash
# round-2 evidence session: plain bench (exit 0) -> ncu launch list of the same command -> ncu --set full of the top kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python vaesne-dev_b200/build.py --source-hash > gpurun_out/r2_ncu_src.sha256
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-profile > gpurun_out/r2_ncu_bench.log 2>&1
for k in attn_tc_fwdN attn_tc_bwd1; do
  CHECK_BWD=1 ONLY_TIME=1 PS=0.1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_$k \
    python tests/probe/attn_tc_check.py > gpurun_out/r2_ncu_$k.log 2>&1
done
for k in lin_tc_bwd lin_tc_fwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2_${k}_ln \
    python tests/probe/lin_bench.py > gpurun_out/r2_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -6
