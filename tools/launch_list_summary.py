"""launch_list_summary.py LIST.csv OUT.json "command" "note" -- per-kernel launches / time / share of an
`ncu --metrics gpu__time_duration.sum --csv` launch list (serialised, cold-cache per-launch times: compare SHARES)."""
import csv
import json
import re
import sys


def main():
    src, dst, command, note = sys.argv[1:5]
    rows, hdr = [], None
    with open(src) as f:
        for r in csv.reader(f):
            if hdr is None:
                if "Kernel Name" in r:
                    hdr = r
                continue
            rows.append(dict(zip(hdr, r)))
    kern, total = {}, 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        ns = float(r["Metric Value"].replace(",", ""))
        ns *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
        k = kern.setdefault(name, {"launches": 0, "ms": 0.0})
        k["launches"] += 1
        k["ms"] += ns / 1e6
        total += ns / 1e6
    for k in kern.values():
        k["share"] = k["ms"] / total if total else 0.0
    out = {"round": 2, "command": command, "note": note, "launches_in_window": len(rows), "total_ms": total,
           "kernels": dict(sorted(kern.items(), key=lambda kv: -kv[1]["ms"]))}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: (v["launches"], round(v["share"], 3)) for k, v in out["kernels"].items()}))


if __name__ == "__main__":
    main()
