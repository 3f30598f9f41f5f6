"""DRAM bytes of the last N kernel launches of an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` log.
Usage: python profiles/tools/traffic_from_csv.py log.csv N  -> prints a JSON object"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0] != "ID"]
n_last = int(sys.argv[2])
ids = sorted({int(r[0]) for r in rows})
keep = set(ids[-n_last:])
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0}
names = {}
for r in rows:
    if int(r[0]) in keep and r[12] in tot:
        tot[r[12]] += float(r[14].replace(",", "")) * mult[r[13]]
        names[int(r[0])] = r[4].split("(")[0].split("::")[-1]
print(json.dumps({"launches": len(keep), "dram_read_bytes": tot["dram__bytes_read.sum"], "dram_write_bytes": tot["dram__bytes_write.sum"],
                  "first_kernel": names[min(keep)], "last_kernel": names[max(keep)]}))
