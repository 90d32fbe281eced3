#!/bin/bash
# round-2 GPU session B: the full bench line (default flags) + the reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
cat gpurun_out/r2b_bench_ref.json; tail -3 gpurun_out/r2b_bench_ref.err
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -5 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "cpu_baseline", "eager_gpu_baseline", "configs", "batch_sweep", "encode"):
    print(k, json.dumps(d.get(k))[:900])
print("roofline", json.dumps(d["roofline"])[:1500])
PY
