timeout 120 python -m pytest tests/test_gpu_metrics.py -m gpu -x -q 2>&1 | tail -8
RC=16 timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc_ring -s 2 -c 1 -f -o gpurun_out/ring16_v12 python tools/_ring_one.py 2>&1 | tail -1
RC=32 timeout 200 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc_ring -s 2 -c 1 -f -o gpurun_out/ring32_v12 python tools/_ring_one.py 2>&1 | tail -1
