"""profiles/traffic_r02.json from an ncu launch list of `bench.py --profile` (one V-cycle, plain launches):

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --cpu-cycles 0 --profile
    python tools/ncu_traffic.py gpurun_out/launches.csv

For every row-op of the fine level (the largest launch of each (LANES, OP) template instance) the DRAM bytes of that
launch.  The record carries the sha256 of csrc/apply.cu: bench.py reports `roofline.traffic` from it only while the
kernel source is unchanged."""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = {0: "OP_SPMV", 1: "OP_SPMV_ADD", 2: "OP_RESIDUAL", 3: "OP_JACOBI", 4: "OP_RESZERO", 5: "OP_PSMOOTH", 6: "OP_RESZERO_S",
       7: "OP_PSMOOTH0"}


def main():
    path = sys.argv[1]
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    launches = {}
    for r in csv.DictReader(lines):
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "bytes": 0.0, "ns": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            d["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif r["Metric Name"] == "gpu__time_duration.sum":
            d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(u, 1)
    best = {}
    for d in launches.values():
        m = re.search(r"csr_rowop_kernel<double, (\d+), (\d+)", d["name"])
        mw = re.search(r"csr_w32_rowop_kernel<double, (\d+)", d["name"])
        if mw:
            key = f"csr_w32_rowop_kernel<double,{OPS.get(int(mw.group(1)), mw.group(1))}>"
        elif m:
            key = f"LANES={int(m.group(1))},{OPS.get(int(m.group(2)), m.group(2))}>"
        else:
            continue
        if key not in best or d["bytes"] > best[key]["bytes"]:
            best[key] = d
    sha = hashlib.sha256(open(os.path.join(ROOT, "ml-amg_b200", "csrc", "apply.cu"), "rb").read()).hexdigest()
    out = {"apply_cu_sha256": sha,
           "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, largest launch of each row-op instance in {os.path.basename(path)} "
                     "(one plain-launch V(1,1) cycle of bench.py --profile at 256^3)",
           "dram_bytes_per_launch": {k: int(v["bytes"]) for k, v in sorted(best.items())},
           "us_per_launch_under_ncu": {k: round(v["ns"] / 1e3, 1) for k, v in sorted(best.items())}}
    with open(os.path.join(ROOT, "profiles", "traffic_r02.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
