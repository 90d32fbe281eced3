#!/bin/bash
# round-2 closing session: full GPU suite, smoke, reference arm, default bench, evidence captures for the final sources
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; cut -c1-300 gpurun_out/r2f_bench_ref.json
( time timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "cpu_baseline", "eager_gpu_baseline", "batch_sweep", "encode"):
    print(k, json.dumps(d.get(k))[:600])
print("roofline", json.dumps(d["roofline"])[:1800])
PY
bash tests/probe/gpu_ncu_r2.sh > gpurun_out/r2f_ncu.log 2>&1; tail -4 gpurun_out/r2f_ncu.log
