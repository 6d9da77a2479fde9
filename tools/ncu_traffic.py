"""profiles/ncu_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every
kernel of the pipeline, from the summaries of the `ncu --set full` captures (tools/gpu_profile.sh,
tools/ncu_summary.py).  bench.py reads it for the `traffic` field of its roofline objects.

  python tools/ncu_traffic.py WORKLOAD EM_STEPS.csv ONCE.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GEMMS = ["gemm_conv1", "gemm_conv2", "gemm_conv3", "gemm_conv4", "gemm_conv5", "gemm_fc6", "gemm_fc7", "gemm_fc8"]
SPLITK = ["splitk_fc6", "splitk_fc7", "splitk_fc8"]


def rows(path):
    r = list(csv.reader(open(path)))
    h, units = r[0], r[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for x in r[2:]:
        d = dict(zip(h, x))
        u = dict(zip(h, units))
        byt = sum(float(d[k]) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tu = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u["gpu__time_duration.sum"]]
        out.append((d["Kernel Name"], byt, float(d["gpu__time_duration.sum"]) * tu, int(d["launch__grid_size"])))
    return out


def main():
    workload, em_csv, once_csv = sys.argv[1:4]
    acc = {}

    def add(name, byt, ms, grid):
        a = acc.setdefault(name, {"bytes": 0.0, "ms": 0.0, "launches": 0, "grid": grid})
        a["bytes"] += byt; a["ms"] += ms; a["launches"] += 1

    for name, byt, ms, grid in rows(em_csv):
        if ms > 0.2 and "post" in name:
            continue            # the superstep in which split_best_vp runs: not the steady-state launch
        add(name.replace("void ", "").split("<")[0].replace("_kernel", ""), byt, ms, grid)
    gi = li = si = 0
    for name, byt, ms, grid in rows(once_csv):
        if "gemm_bf16" in name:
            add(GEMMS[gi], byt, ms, grid); gi += 1
        elif "splitk" in name:
            add(SPLITK[si], byt, ms, grid); si += 1
        elif "lrn_pool_kernel" in name:           # the three max-pools, in launch order (the LRN lives in the conv epilogues)
            add(["pool1", "pool2", "pool5"][min(li, 2)], byt, ms, grid); li += 1
        else:
            add(name.replace("void ", "").split("<")[0].replace("_kernel", ""), byt, ms, grid)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    doc[workload] = {k: {"dram_bytes_per_launch": v["bytes"] / v["launches"], "ncu_ms_per_launch": v["ms"] / v["launches"],
                         "captured_launches": v["launches"], "grid": v["grid"]} for k, v in sorted(acc.items())}
    doc[workload]["_source"] = [os.path.basename(em_csv), os.path.basename(once_csv)]
    doc[workload]["_note"] = ("ncu --set full --clock-control none, cold caches, serialised launches; the EM superstep kernels were "
                              "captured in supersteps 10-12 (all images of the batch still active)")
    json.dump(doc, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, sorted(acc))


if __name__ == "__main__":
    main()
