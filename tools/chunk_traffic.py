"""Sum DRAM traffic and durations over the kernels of an `ncu -i X.ncu-rep --page raw --csv` dump (one activation chunk of the
PPO update, or any launch range):  python tools/chunk_traffic.py raw.csv [out.json] [key=value ...]"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
units = rows[1]


def val(r, name):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return 0.0
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3,
             "msecond": 1e3, "second": 1e6}.get(u, 1)
    return v * scale


kern = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    kern.append(dict(kernel=name, us=val(r, "gpu__time_duration.sum"), dram_read=val(r, "dram__bytes_read.sum"),
                     dram_write=val(r, "dram__bytes_write.sum"),
                     tensor_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
# optional window: first=<kernel-name substring> count=<n> keeps n kernels starting at the first match (one super-chunk of the update)
opts = dict(kv.split("=", 1) for kv in sys.argv[2:] if "=" in kv)
if "first" in opts:
    i0 = next(i for i, k in enumerate(kern) if opts["first"] in k["kernel"])
    kern = kern[i0:i0 + int(opts.get("count", len(kern)))]
tot_b = sum(k["dram_read"] + k["dram_write"] for k in kern)
tot_us = sum(k["us"] for k in kern)
for k in kern:
    print("%-58s %8.1f us  r %8.1f MB  w %8.1f MB  dram %5.1f%%  tensor %5.1f%%" % (k["kernel"][:58], k["us"], k["dram_read"] / 1e6,
                                                                                 k["dram_write"] / 1e6, k["dram_pct"], k["tensor_pct"]))
print("total %d kernels, %.1f us, %.1f MB DRAM (%.2f TB/s over the summed durations)" % (len(kern), tot_us, tot_b / 1e6, tot_b / tot_us / 1e6))
if len(sys.argv) > 2 and not "=" in sys.argv[2]:
    out = dict(dram_bytes_per_chunk=tot_b, sum_kernel_us=tot_us, kernels=kern)
    for kv in sys.argv[3:]:
        k, v = kv.split("=", 1)
        if k in ("first", "count"):
            continue
        try:
            v = json.loads(v)
        except Exception:
            pass
        out[k] = v
    json.dump(out, open(sys.argv[2], "w"), indent=1)
