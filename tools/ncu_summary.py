"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list by
kernel: launches, total / mean duration, DRAM bytes, achieved DRAM GB/s and its fraction of the measured copy peak.

    python tools/ncu_summary.py gpurun_out/launches.csv [--peak 6542.1] [--top 25] > profiles/r02_....md"""
import csv
import collections
import re
import sys


def main():
    path = sys.argv[1]
    peak = float(sys.argv[sys.argv.index("--peak") + 1]) if "--peak" in sys.argv else 6542.1
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    acc = collections.OrderedDict()
    for r in rd:
        name = r.get("Kernel Name", "")
        name = re.sub(r"\(.*$", "", name)
        key = (r.get("ID"), name)
        d = acc.setdefault(key, {"name": name})
        m, v, u = r.get("Metric Name"), r.get("Metric Value", "0").replace(",", ""), r.get("Metric Unit", "")
        try:
            v = float(v)
        except ValueError:
            continue
        if m == "gpu__time_duration.sum":
            d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1, "KB": 1e3, "MB": 1e6, "GB": 1e9}.get(u, 1)
            d["bytes"] = d.get("bytes", 0.0) + v * mult
    by = collections.OrderedDict()
    for d in acc.values():
        k = by.setdefault(d["name"], {"n": 0, "ns": 0.0, "bytes": 0.0})
        k["n"] += 1
        k["ns"] += d.get("ns", 0.0)
        k["bytes"] += d.get("bytes", 0.0)
    total = sum(k["ns"] for k in by.values())
    print(f"launches: {sum(k['n'] for k in by.values())}, total kernel time {total / 1e6:.2f} ms (ncu: serialised, cold caches)\n")
    print("| kernel | launches | total ms | share | mean us | DRAM GB moved | GB/s | of peak |")
    print("|---|---|---|---|---|---|---|---|")
    for name, k in sorted(by.items(), key=lambda kv: -kv[1]["ns"])[:top]:
        gbs = k["bytes"] / k["ns"] if k["ns"] else 0.0
        print(f"| `{name[:90]}` | {k['n']} | {k['ns'] / 1e6:.2f} | {100 * k['ns'] / total:.1f}% | {k['ns'] / 1e3 / k['n']:.1f} | "
              f"{k['bytes'] / 1e9:.2f} | {gbs:.0f} | {gbs / peak:.2f} |")


if __name__ == "__main__":
    main()
