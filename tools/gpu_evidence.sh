#!/usr/bin/env bash
# Evidence bundle at HEAD (one gpurun call, ~4 min): per-kernel DRAM traffic of one step, the launch list of the bench
# command, ncu --set full captures of the kernels DESIGN.md discusses.  Afterwards, here:
#   python tools/traffic_table.py gpurun_out/step.csv gpurun_out/tags.json profiles/traffic_rNN.json
#   python tools/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/ncu_<name>_rNN.txt
set -u
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/step.csv python tools/profile_step.py --steps 1 --tags gpurun_out/tags.json 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cmd.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-full-codd > gpurun_out/b.log 2>&1
for c in 16 32; do
  RC=$c timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc_ring -s 2 -c 1 -f \
      -o gpurun_out/ring$c python tools/ring_one.py 2>&1 | tail -1
done
timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv3x3x2 -s 2 -c 1 -f \
    -o gpurun_out/ring2 python tools/ring2_one.py 2>&1 | tail -1
K4_KINDS=init timeout 300 ncu --set full --import-source on --clock-control none -k regex:tile_warp_cost -s 3 -c 1 -f \
    -o gpurun_out/k4 python tools/k4_probe.py 2>&1 | tail -1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:se3_gn_step -s 2 -c 1 -f \
    -o gpurun_out/se3_gn python tools/bench_full_codd.py 2>&1 | tail -1
timeout 60 python tools/ring2_one.py; timeout 60 python tools/upmerge_one.py
