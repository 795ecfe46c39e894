#!/usr/bin/env bash
# One gpurun bundle for a kernel change: the ops unit tests, the end-to-end stereo parity tests, a short bench with the
# per-kernel table.   gpurun --timeout 900 -- 'bash tools/gpu_step.sh'
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_hitnet.py tests/test_gpu_parity_headline.py -m gpu -q --tb=short 2>&1 | tail -6
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-full-codd \
    --dump-kernels gpurun_out/kernels_step.json > gpurun_out/bench_step.json 2> gpurun_out/bench_step.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_step.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for r in json.load(open("gpurun_out/kernels_step.json"))[:24]:
    print("%-34s %2d x  %.4f ms  %.3f" % (r["kernel"], r["launches_per_step"], r["ms_per_step"], r["frac"]))
PY
