"""Per-kernel share table of one bench step from a PARTIAL ncu launch list (--launch-skip N -c M over the tail of the step: the last
ADMM iterations + the matching of all slices).  Iterations are identical, so the step is rebuilt as 100 x-updates + 99 denoiser passes
+ one matching per slice.  Usage: python profiles/tools/launch_table_tail.py gpurun_out/r02y_launches_tail.csv [slices] > profiles/<name>.md"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 120
mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "s": 1e6}
def name(r):
    return re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
seq = [(name(r), float(r[14].replace(",", "")) * mult.get(r[13], 1.0)) for r in rows]
starts = [i for i, (n, _) in enumerate(seq) if n.startswith("stream_fwd_kernel")]
first_match = next(i for i, (n, _) in enumerate(seq) if n.startswith("match_prep"))
full = [(a, b) for a, b in zip(starts[:-1], starts[1:])]           # complete iterations: x-update + denoiser pass
it = collections.OrderedDict()
for a, b in full:
    for n, us in seq[a:b]:
        e = it.setdefault(n, [0, 0.0]); e[0] += 1; e[1] += us
nit = len(full)
match = collections.OrderedDict()
nm = 0
for n, us in seq[first_match:]:
    if not n.startswith("match"): break
    e = match.setdefault(n, [0, 0.0]); e[0] += 1; e[1] += us
    nm += n.startswith("match_prep")
xu = {n: v for n, v in it.items() if n.startswith("stream_") or n.startswith("minmax")}
den = {n: v for n, v in it.items() if n not in xu}
tot_x = sum(v[1] for v in xu.values()) / nit
tot_d = sum(v[1] for v in den.values()) / nit
tot_m = sum(v[1] for v in match.values()) / nm
step = 100 * tot_x + 99 * tot_d + S * tot_m
print(f"# Launch list of the tail of ONE step of the {S}-slice job (BASELINE configs[3]) - B200, end of round 2\n")
print(f"`ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6300 -c 1350 --csv python bench.py --steps 1 --warmup 0 --skip-cpu --skip-extra`")
print(f"(`{sys.argv[1].split('/')[-1]}`: {len(seq)} launches = the last {nit} complete PnP-ADMM iterations, the final x-update and the matching of {nm} slices; the")
print("GPU-minute budget did not allow all ~7300 launches of the step again).  Iterations are identical, so the step is rebuilt as 100 x-updates +")
print(f"99 denoiser passes + {S} matchings.  Per-launch times under ncu are cold-cache, serialised and taken at burst clocks: compare SHARES.\n")
print("| kernel | launches per step | us per launch | us per step | share |")
print("|---|---|---|---|---|")
def line(n, per_unit, us_unit, units):
    print(f"| `{n}` | {per_unit * units:.0f} | {us_unit / per_unit:.1f} | {us_unit * units:.0f} | {100 * us_unit * units / step:.2f} % |")
allk = [(n, v[0] / nit, v[1] / nit, 99) for n, v in den.items()] + [(n, v[0] / nit, v[1] / nit, 100) for n, v in xu.items()] + \
       [(n, v[0] / nm, v[1] / nm, S) for n, v in match.items()]
for n, c, us, units in sorted(allk, key=lambda t: -t[2] * t[3]):
    line(n, c, us, units)
print(f"| **one step** | {sum(c * u for _, c, _, u in allk):.0f} | | {step:.0f} | 100 % |\n")
print(f"Groups: denoiser {100 * 99 * tot_d / step:.2f} % ({tot_d / 1e3:.2f} ms per pass), matching {100 * S * tot_m / step:.2f} % ({tot_m / 1e3:.3f} ms per slice), "
      f"x-update {100 * 100 * tot_x / step:.2f} % ({tot_x / 1e3:.3f} ms per {S}-slice update).")
