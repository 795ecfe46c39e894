#!/usr/bin/env bash
# Round-2 first GPU bundle: full -m gpu suite (no -x: every failure is reported), smoke, default bench (with the
# unmodified-reference CPU / GPU-eager baselines and the full-CODD leg), the staged 1x1 experiment (DIAG build only).
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | head -25 >> gpurun_out/topo.txt; nproc >> gpurun_out/topo.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | tail -150 > gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e_f32"]["value"], d.get("e2e_graph", {}).get("value"))
print("gpu eager", d.get("gpu_eager_baseline")); print("cpu", d.get("cpu_baseline")); print("full", d.get("full_codd"))
print([(k["kernel"], k["ms"]) for k in d["top_kernels"]])
print([(k["kernel"], k["frac"], k.get("ms_per_step", k.get("ms_per_launch"))) for k in d["roofline_named_kernels"]])
PY
if [ "${1:-}" = "staged" ]; then
  CODD_PW_STAGED=1 timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "conv" 2>&1 | tail -3
  for v in 0 1; do
    CODD_PW_STAGED=$v timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-full-codd 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('staged=$v', d['value'], d['ms_per_step'], [(k['kernel'],k['ms']) for k in d['top_kernels'] if '1x1' in k['kernel']])"
  done
fi
