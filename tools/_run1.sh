timeout 100 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "ring" 2>&1 | tail -3
timeout 100 python tools/ring_check.py 2>&1 | tail -5
timeout 60 python tools/conv_probe.py 2>&1 | grep "N="| sed 's/halo-tile.*ring/ring/'
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_dec.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value']);print([(k['kernel'],k['ms']) for k in d['top_kernels']])"
