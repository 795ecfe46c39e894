#!/usr/bin/env bash
# K4 bundle: bit-exact tests, end-to-end tests, probe timings, short bench, one ncu --set full capture of the kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_hitnet.py tests/test_gpu_parity_headline.py tests/test_gpu_boundary.py -m gpu -q --tb=short 2>&1 | tail -25
timeout 120 python tools/k4_probe.py 2>&1 | tail -4
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-full-codd > gpurun_out/bench_k4.json 2> gpurun_out/bench_k4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_k4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
print([(k["kernel"], k["ms"]) for k in d["top_kernels"]])
print([(k["kernel"], k["frac"], k.get("ms_per_step", k.get("ms_per_launch"))) for k in d["roofline_named_kernels"]])
PY
if [ "${1:-}" = "ncu" ]; then
  K4_KINDS=init timeout 300 ncu --set full --import-source on --clock-control none -k regex:tile_warp_cost -s 3 -c 1 -f \
      -o gpurun_out/k4_r02 python tools/k4_probe.py 2>&1 | tail -2
fi
