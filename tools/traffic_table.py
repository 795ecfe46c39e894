#!/usr/bin/env python
"""Join an ncu launch list (csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per
launch) of `tools/profile_step.py --steps 1 --tags tags.json` with the ops tag sequence of that step, and write
the per-kernel-tag table bench.py reads for `roofline.traffic`:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/step.csv python tools/profile_step.py --steps 1 --tags gpurun_out/tags.json
    python tools/traffic_table.py gpurun_out/step.csv gpurun_out/tags.json profiles/traffic_r01.json
"""
import csv
import json
import sys


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(value, unit):
    v = float(value.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit.lower(), 1)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    idx = {h: i for i, h in enumerate(rows[start])}
    launches = {}
    for r in rows[start + 1:]:
        if len(r) <= idx["Metric Value"] or not r[0].isdigit():
            continue
        d = launches.setdefault(int(r[0]), dict(kernel=r[idx["Kernel Name"]]))
        name, unit, val = r[idx["Metric Name"]], r[idx["Metric Unit"]], r[idx["Metric Value"]]
        if name == "gpu__time_duration.sum":
            d["us"] = to_us(val, unit)
        elif name == "dram__bytes_read.sum":
            d["rd"] = to_bytes(val, unit)
        elif name == "dram__bytes_write.sum":
            d["wr"] = to_bytes(val, unit)
    tags = json.load(open(sys.argv[2]))
    # the library's kernels live in an anonymous namespace; torch's own launches (input synthesis, weight init)
    # are skipped, and the tagged step is the LAST one profiled
    seq = [launches[k] for k in sorted(launches) if launches[k]["kernel"].startswith(("void <unnamed>::", "<unnamed>::"))][-len(tags):]
    assert len(seq) == len(tags), f"{len(seq)} profiled launches vs {len(tags)} tagged ops"
    table = {}
    for (tag, nbytes), l in zip(tags, seq):
        t = table.setdefault(tag, dict(kernel=l["kernel"].replace("void <unnamed>::", "")[:60], launches=0, us=0.0,
                                       dram_read=0.0, dram_write=0.0, algorithmic=0))
        t["launches"] += 1
        t["us"] += l["us"]
        t["dram_read"] += l.get("rd", 0.0)
        t["dram_write"] += l.get("wr", 0.0)
        t["algorithmic"] += nbytes
    for t in table.values():
        n = t["launches"]
        t["traffic_per_launch"] = (t["dram_read"] + t["dram_write"]) / n
        t["algorithmic_per_launch"] = t["algorithmic"] / n
        t["us_per_launch_under_ncu"] = t["us"] / n
    total = sum(t["us"] for t in table.values())
    for t in table.values():
        t["share_of_step_under_ncu"] = t["us"] / total
    json.dump(dict(source=sys.argv[1], note="one eager step of tools/profile_step.py under ncu (cold caches, serialised); "
                   "dram bytes = dram__bytes_read.sum + dram__bytes_write.sum", kernels=table), open(sys.argv[3], "w"), indent=1)
    for tag, t in sorted(table.items(), key=lambda kv: -kv[1]["us"])[:12]:
        print(f"{tag:40s} x{t['launches']:3d} {t['us']:9.1f} us  share {t['share_of_step_under_ncu']:.3f}  "
              f"traffic/alg {t['traffic_per_launch'] / max(t['algorithmic_per_launch'], 1):.2f}")


if __name__ == "__main__":
    main()
