timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e_u8']['value'],d['e2e_graph']['value'])"; tail -3 gpurun_out/bench_final.err
