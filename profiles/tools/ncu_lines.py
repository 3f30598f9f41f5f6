"""Per-source-line share of executed instructions, stall samples and shared-memory wavefronts of one kernel.
Usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > src.csv ; python profiles/tools/ncu_lines.py src.csv [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
def num(x):
    try: return float(x)
    except Exception: return 0.0
sec, hdr, data = None, None, {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': sec = r[1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or sec is None: continue
    try: ln = int(r[0])
    except Exception: continue
    d = dict(zip(hdr, r))
    key = (sec.split('/')[-1], ln)
    ie, st = num(d.get('Instructions Executed')), num(d.get('Warp Stall Sampling (All Samples)'))
    wf, wfi = num(d.get('L1 Wavefronts Shared')), num(d.get('L1 Wavefronts Shared Ideal'))
    if ie or st:
        a = data.setdefault(key, [0, 0, 0, 0, d.get('Source', '')[:80]])
        a[0] += ie; a[1] += st; a[2] += wf; a[3] += wfi
tot = sum(v[0] for v in data.values()); tots = sum(v[1] for v in data.values())
print('total warp instructions', tot, 'stall samples', tots)
for k, v in sorted(data.items()):
    if v[0] > thr * tot or v[1] > thr * tots:
        print(f"{k[0]:22s} {k[1]:4d} inst {100*v[0]/tot:5.1f}% stall {100*v[1]/tots:5.1f}% smem wavefronts {v[2]/1e6:6.2f}M (ideal {v[3]/1e6:6.2f}M) | {v[4].strip()}")
