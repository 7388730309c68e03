"""profiles/*_traffic_per_kernel.json from an ncu csv log:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>
   python scripts/ncu_traffic.py X.csv "<description>" > profiles/rN_traffic_per_kernel.json"""
import collections, csv, json, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H = rows[hdr]
ik, im, iu, iv, iid = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Unit"), H.index("Metric Value"), H.index("ID")
per = collections.defaultdict(lambda: collections.defaultdict(float))
launches = collections.defaultdict(set)
for r in rows[hdr + 1:]:
    name = r[ik].split("(")[0].replace("void ", "")
    per[name][r[im]] += float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
    launches[name].add(r[iid])
out = {"source": sys.argv[2] if len(sys.argv) > 2 else "", "kernels": {}}
for k, m in sorted(per.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
    n = len(launches[k])
    rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
    out["kernels"][k] = {"launches": n, "dram_read_bytes_total": rd, "dram_write_bytes_total": wr, "dram_bytes_per_launch": (rd + wr) / max(1, n),
                         "l2_bytes_total": m.get("lts__t_bytes.sum", 0.0), "time_s_total": m.get("gpu__time_duration.sum", 0.0)}
print(json.dumps(out, indent=1))
