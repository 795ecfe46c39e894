timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 200 python bench.py --no-cpu-baseline --dump-kernels gpurun_out/kernels_final.json --steps 10 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cmd.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1; tail -2 gpurun_out/b.log | cut -c1-200
