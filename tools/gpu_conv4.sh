#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "conv4x4_stride2 or tile_features_tensor or upmerge" -s 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_hitnet.py tests/test_gpu_parity_headline.py -m gpu -q --tb=short 2>&1 | tail -8
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-full-codd > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
print([(k["kernel"], k["ms"]) for k in d["top_kernels"]])
PY
