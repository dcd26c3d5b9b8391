"""ncu full capture(s) -> profiles/r02_traffic.json (what bench.py reports as roofline.traffic) + a per-launch summary.
    python profiles/make_traffic.py fp32=gpurun_out/r02_prof_fp32_gin.ncu-rep bf16=gpurun_out/r02_prof_bf16_gin.ncu-rep
Reads the reports with `ncu -i ... --page raw --csv` (no GPU needed).  A roofline unit is ONE GIN layer of both encoders:
forward = one launch, backward = gin_bwd_pre + gin_bwd_main (layer average over the captured launches)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def load(rep):
    if rep.endswith(".csv"):     # already exported on the GPU box (`ncu -i rep --page raw --csv`): the reports exceed gpurun's copy-back cap
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]]}
        for k in KEEP:
            if k in idx:
                v = float(r[idx[k]].replace(",", ""))
                d[k] = v * SCALE.get(units[idx[k]], 1.0)        # bytes / microseconds / plain
        res.append(d)
    return res


def family(name):
    if "gin_fwd" in name:
        return "gin_fwd"
    if "gin_bwd_pre" in name:
        return "gin_bwd_pre"
    if "gin_bwd_h_kernel<1>" in name or "gin_bwd_h_kernel<true>" in name:
        return "gate_lin_bwd_h"          # half mode: the compressor's linear layer
    if "gin_bwd" in name:
        return "gin_bwd_main"
    return name.split("(")[0].split("<")[0].split("::")[-1]


def main():
    traffic, summary = {}, {}
    for arg in sys.argv[1:]:
        dtype, rep = arg.split("=")
        launches = load(os.path.join(ROOT, rep) if not os.path.isabs(rep) else rep)
        fam = {}
        for l in launches:
            f = fam.setdefault(family(l["kernel"]), [])
            f.append(l)
        # the head MLP backward runs on the same kernel as the GIN layers (N instead of N + Ns rows): not a roofline unit
        if "gin_bwd_main" in fam:
            big = max(l["dram__bytes_read.sum"] for l in fam["gin_bwd_main"])
            head = [l for l in fam["gin_bwd_main"] if l["dram__bytes_read.sum"] < 0.5 * big]
            if head:
                fam["head_bwd"] = head
                fam["gin_bwd_main"] = [l for l in fam["gin_bwd_main"] if l["dram__bytes_read.sum"] >= 0.5 * big]
        summ = {}
        for f, ls in fam.items():
            n = len(ls)
            summ[f] = {"launches": n}
            for k in KEEP:
                if k in ls[0]:
                    summ[f][k + " (mean)"] = sum(l[k] for l in ls) / n
            summ[f]["dram_bytes_per_launch"] = summ[f]["dram__bytes_read.sum (mean)"] + summ[f]["dram__bytes_write.sum (mean)"]
        summary[dtype] = summ
        t = {}
        fwd_name = "gin_fwd_bf16.enc1+2" if dtype == "bf16" else "gin_fwd_tc.enc1+2"
        if "gin_fwd" in summ:
            t[fwd_name] = {"dram_bytes_per_unit": summ["gin_fwd"]["dram_bytes_per_launch"]}
        if "gin_bwd_pre" in summ and "gin_bwd_main" in summ:
            t["gin_bwd (pre + main)"] = {"dram_bytes_per_unit": summ["gin_bwd_pre"]["dram_bytes_per_launch"] + summ["gin_bwd_main"]["dram_bytes_per_launch"],
                                         "pre": summ["gin_bwd_pre"]["dram_bytes_per_launch"], "main": summ["gin_bwd_main"]["dram_bytes_per_launch"]}
        traffic[dtype] = t
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    json.dump(summary, open(os.path.join(ROOT, "profiles", "r02_ncu_full_summary.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
