"""Generates tests/golden/metrics_golden.json from the REFERENCE functions (utils/misc.py, utils/warp.py,
utils/metric.py driven as in model/codd.py:462-515).  Run in the build container: python tests/golden/gen_metrics_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_loader  # noqa: E402
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "test_metrics_oracle", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "test_metrics_oracle.py"))
_t = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_t)
make_case, reference_metrics = _t.make_case, _t.reference_metrics

U = ref_loader.load().utils
out = []
for seed, kitti in [(10, False), (11, False), (12, True)]:
    r = reference_metrics(U, make_case(seed, kitti=kitti))
    r.pop("mask_disp"); r.pop("warp")
    out.append(dict(seed=seed, kitti=kitti, ref=r))
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.json"), "w"), indent=1)
print(out)
