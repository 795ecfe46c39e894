#!/usr/bin/env bash
# Standard GPU bundles for `gpurun` (one call each), with tight per-command timeouts: a hung kernel otherwise burns the
# whole call limit (and a dead GPU is a strike).  Charged time is ~20 s of overhead + the run time.
#   gpurun --timeout 150 -- 'bash tools/gpu_check.sh quick'     ring / conv unit tests + conv probe            (~15 s run)
#   gpurun --timeout 400 -- 'bash tools/gpu_check.sh full'      full -m gpu suite, smoke(), default bench      (~60 s run)
#   gpurun --timeout 600 -- 'bash tools/gpu_check.sh profile'   step traffic table + launch list + ring ncu    (~150 s run)
set -u
mkdir -p gpurun_out
case "${1:-quick}" in
  quick)
    timeout 100 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "ring or tc or conv" 2>&1 | tail -3
    timeout 60 python tools/conv_probe.py 2>&1 | grep "N=" ;;
  full)
    timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
    timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
    timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
    python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("e2e_u8", {}).get("value"), d.get("e2e_graph"))
print([(k["kernel"], k["ms"]) for k in d["top_kernels"]])
PY
    ;;
  profile)
    timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/step.csv python tools/profile_step.py --steps 1 --tags gpurun_out/tags.json 2>&1 | tail -1
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cmd.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
    for c in 16 32; do
      RC=$c timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc_ring -s 2 -c 1 -f \
          -o gpurun_out/ring$c python tools/ring_one.py 2>&1 | tail -1
    done
    echo "then here: python tools/traffic_table.py gpurun_out/step.csv gpurun_out/tags.json profiles/traffic_rNN.json" ;;
  *) echo "usage: $0 quick|full|profile"; exit 2 ;;
esac
