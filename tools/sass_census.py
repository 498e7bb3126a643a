"""sass_census.py -- per-kernel SASS mnemonic census of gpax_b200/lib/libb200gp.so (cuobjdump -sass): which kernels carry the
Blackwell-native instructions (UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,
UCGABAR = cluster barrier, SYNCS = mbarrier, DMMA = fp64 tensor MMA).  Writes profiles/sass_census_r2.json."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpax_b200", "lib", "libb200gp.so")
MNEMONICS = ["UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UCGABAR", "SYNCS", "DMMA", "HMMA", "IMMA", "LDGSTS", "MUFU", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    census, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            census[cur] = {k: 0 for k in MNEMONICS}
            census[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1).split(".")[0]
        census[cur]["instructions"] += 1
        for k in MNEMONICS:
            if op.startswith(k):
                census[cur][k] += 1
    rows = {k: {m: v for m, v in c.items() if v} for k, c in census.items()}
    total = {m: sum(c.get(m, 0) for c in census.values()) for m in MNEMONICS}
    res = {"library": os.path.relpath(LIB, ROOT), "command": "cuobjdump -sass gpax_b200/lib/libb200gp.so", "totals": total, "kernels": rows}
    path = os.path.join(ROOT, "profiles", "sass_census_r2.json")
    json.dump(res, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(total))
    for k in sorted(rows):
        if any(m in rows[k] for m in ("UTCIMMA", "UTMALDG", "DMMA", "LDTM")):
            print(k, {m: rows[k][m] for m in ("UTCIMMA", "LDTM", "UTMALDG", "UTCBAR", "UCGABAR", "DMMA") if m in rows[k]})


if __name__ == "__main__":
    main()
