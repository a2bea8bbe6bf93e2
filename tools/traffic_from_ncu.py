#!/usr/bin/env python
"""profiles/r02_traffic_k_hmm.json from an ncu launch list of tools/stage_bench.py taken with
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_hmm --csv
(tools/gpu_visit.sh step `traffic`): DRAM bytes of the HMM launch set of ONE step (the last one captured), which
bench.py scales by band cells into `roofline.traffic`.

  python tools/traffic_from_ncu.py gpurun_out/<tag>_traffic.csv gpurun_out/<tag>_traffic_stage.json profiles/r02_traffic_k_hmm.json
"""
import csv
import json
import sys


def main():
    src, stage_json, dst = sys.argv[1:4]
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    ki, mi, vi, ui = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
    launches = {}
    for r in rows[h + 1:]:
        if len(r) <= vi or not r[0].isdigit():
            continue
        nm = r[ki]
        nm = nm[:nm.index(">(") + 1] if ">(" in nm else nm.split("(")[0]
        d = launches.setdefault(int(r[0]), {"kernel": nm.replace("(int)", "").replace("(bool)", "")})
        v = float(r[vi].replace(",", ""))
        unit = r[ui].lower()
        if "byte" in unit:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        d[r[mi]] = v
    ids = sorted(launches)
    # every step launches the same set: steps = how often the most common kernel of the bulk class appears
    names = [launches[i]["kernel"] for i in ids]
    bulk = max(set(names), key=lambda n: (("41" in n), names.count(n)))
    steps = max(1, names.count(bulk))
    per_step = len(names) // steps
    last = ids[-per_step:]
    st = json.loads(open(stage_json).read().strip().splitlines()[-1])
    out = {
        "source": f"{src}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (one step of tools/stage_bench.py "
                  f"--preset {st['preset']} --groups {st['groups']}, {per_step} HMM launches)",
        "workload": st["preset"], "hmm_mode": st["hmm_mode"], "band_cells": st["cells"],
        "dram_bytes_read": sum(launches[i].get("dram__bytes_read.sum", 0) for i in last),
        "dram_bytes_write": sum(launches[i].get("dram__bytes_write.sum", 0) for i in last),
        "launches": [{"kernel": launches[i]["kernel"], "ms": launches[i].get("gpu__time_duration.sum", 0) / 1e6,
                      "dram_bytes": launches[i].get("dram__bytes_read.sum", 0) + launches[i].get("dram__bytes_write.sum", 0)}
                     for i in last],
    }
    out["dram_bytes_total"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    out["bytes_per_cell"] = out["dram_bytes_total"] / out["band_cells"]
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("dram_bytes_total", "bytes_per_cell", "band_cells", "hmm_mode")}))


if __name__ == "__main__":
    main()
