"""ncu per-launch metrics CSV of the conv/GEMM kernel (one forward, bench shape) -> table + JSON summary.
usage: conv_metrics_table.py <csv> <out.json>"""
import csv, collections, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r); h = rows[hi]
idi, mi, vi, ui = h.index("ID"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > vi:
        d.setdefault(int(r[idi]), {})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
NAMES = ["layer1.0.conv1", "layer1.0.conv2", "layer1.1.conv1", "layer1.1.conv2", "layer2.0.conv1", "layer2.0.down",
         "layer2.0.conv2", "layer2.1.conv1", "layer2.1.conv2", "layer3.0.conv1", "layer3.0.down", "layer3.0.conv2",
         "layer3.1.conv1", "layer3.1.conv2", "layer3_outconv", "layer2_outconv", "layer2_outconv2.0", "layer2_outconv2.3",
         "layer1_outconv", "layer1_outconv2.0", "layer1_outconv2.3"]
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def b(x): return x[0] * SC[x[1]]
ids = sorted(d)
out = {"launches": []}
tot = collections.Counter()
print(f"{'launch':22s} {'us':>8s} {'DRAM rd MB':>10s} {'DRAM wr MB':>10s} {'L2 MB':>9s} {'tensor %':>8s} {'SM GHz':>7s}")
for k, i in enumerate(ids):
    m = d[i]
    t = m["gpu__time_duration.sum"]; us = t[0] / 1e3 if t[1].startswith("n") else t[0]
    name = NAMES[k] if k < len(NAMES) else "token / correlation / fine gemm"
    rec = {"name": name, "us": round(us, 1), "dram_read": b(m["dram__bytes_read.sum"]), "dram_write": b(m["dram__bytes_write.sum"]),
           "l2_bytes": b(m["lts__t_bytes.sum"]), "tensor_active_pct": m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0],
           "sm_ghz": round(m["sm__cycles_elapsed.avg.per_second"][0] / 1e9, 3)}
    out["launches"].append(rec)
    grp = "backbone" if k < len(NAMES) else "tokens"
    for key in ("us", "dram_read", "dram_write", "l2_bytes"): tot[grp, key] += rec[key]
    if k < len(NAMES):
        print(f"{name:22s} {us:8.1f} {rec['dram_read']/1e6:10.1f} {rec['dram_write']/1e6:10.1f} {rec['l2_bytes']/1e6:9.1f} {rec['tensor_active_pct']:8.1f} {rec['sm_ghz']:7.3f}")
for grp in ("backbone", "tokens"):
    out[grp] = {k: tot[grp, k] for k in ("us", "dram_read", "dram_write", "l2_bytes")}
    print(f"{grp:22s} {tot[grp,'us']:8.1f} {tot[grp,'dram_read']/1e6:10.1f} {tot[grp,'dram_write']/1e6:10.1f} {tot[grp,'l2_bytes']/1e6:9.1f}")
json.dump(out, open(sys.argv[2], "w"), indent=1)
