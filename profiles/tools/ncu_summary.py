"""Summarise an .ncu-rep (ncu --set full capture) per kernel type: launches, time, DRAM traffic, pipe utilisation.
Usage (build container, no GPU needed): python profiles/tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv, io, subprocess, sys, collections, re

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
def num(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return float("nan")
def scale(k):  # to bytes / microseconds
    u = units[col[k]].lower()
    return {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1
    a["us"] += num(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum")
    a["rd"] += num(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    a["wr"] += num(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    for k, key in [("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   ("lts", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                   ("l1", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
                   ("regs", "launch__registers_per_thread"),
                   ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active")]:
        if key in col:
            v = num(r, key)
            if v == v:
                a[k] += v * (num(r, "gpu__time_duration.sum") if k != "regs" else 1.0)
                a[k + "_w"] += (num(r, "gpu__time_duration.sum") if k != "regs" else 1.0)
tot = sum(a["us"] for a in agg.values())
print(f"ncu --set full --clock-control none, report `{rep.split('/')[-1]}`: {int(sum(a['n'] for a in agg.values()))} launches, {tot:.1f} us summed kernel time "
      "(cold-cache, serialised: compare shares, not absolutes)\n")
print("| kernel | launches | time us | share | avg us | DRAM read MB | DRAM write MB | tensor pipe active % | issue active % | L2 thr % | L1/smem thr % | warps active % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
def w(a, k):
    return a[k] / a[k + "_w"] if a.get(k + "_w") else float("nan")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"| `{name}` | {int(a['n'])} | {a['us']:.1f} | {100*a['us']/tot:.1f}% | {a['us']/a['n']:.2f} | {a['rd']/1e6:.1f} | {a['wr']/1e6:.1f} | "
          f"{w(a,'tensor'):.1f} | {w(a,'issue'):.1f} | {w(a,'lts'):.1f} | {w(a,'l1'):.1f} | {w(a,'warps'):.1f} | {w(a,'regs'):.0f} |")
print(f"\nTotal DRAM traffic: read {sum(a['rd'] for a in agg.values())/1e6:.1f} MB, write {sum(a['wr'] for a in agg.values())/1e6:.1f} MB")
