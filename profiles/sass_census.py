"""Per-kernel SASS census of the shipped library: counts of the mnemonics that prove the Blackwell-native path
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA 1-D), LDGSTS = cp.async) next to FFMA / HMMA.
    python profiles/sass_census.py > profiles/r02_sass_census.json
Runs without a GPU (cuobjdump -sass on s-cgib_b200/lib/libscgib.so)."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "s-cgib_b200", "lib", "libscgib.so")
PAT = collections.OrderedDict([("UTCMMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"),
                               ("UTMALDG", r"\bUTMALDG"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("FFMA", r"\bFFMA"),
                               ("HMMA", r"\bHMMA"), ("F2FP_BF16", r"\bF2FP\.BF16"), ("ACQBULK_PDL", r"\bACQBULK|\bPREEXIT")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("scgib::", "")
            cur = res.setdefault(name, collections.OrderedDict((k, 0) for k in PAT))
            cur["instructions"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"^\s+/\*[0-9a-f]{4,6}\*/", line):
            cur["instructions"] += 1
            for k, p in PAT.items():
                if re.search(p, line):
                    cur[k] += 1
    tot = collections.OrderedDict((k, sum(v[k] for v in res.values())) for k in list(PAT) + ["instructions"])
    json.dump({"library": os.path.relpath(LIB, ROOT), "total": tot,
               "kernels": {k: {a: b for a, b in v.items() if b} for k, v in res.items()}}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
