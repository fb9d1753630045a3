"""Top stalled SASS instructions from `ncu -i rep --page source --csv` output (one section per profiled kernel)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sections, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Address":
        hdr = r
        cur["hdr"] = r
    elif cur is not None and hdr is not None and len(r) == len(hdr):
        cur["rows"].append(r)
for sec in sections:
    hdr = sec["hdr"]; ix = {h: i for i, h in enumerate(hdr)}
    data = sec["rows"]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("==== %s: %d samples, %d instructions" % (sec["name"][:60], tot, len(data)))
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
    print("   stall mix:", ", ".join("%s %.0f%%" % (h[6:], 100 * v / max(tot, 1)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
        s = int(r[ix["# Samples"]])
        st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print("%7d %5.1f%%  %s  %-64s %s" % (s, 100 * s / max(tot, 1), r[ix["Address"]][-5:], r[ix["Source"]].strip()[:64], st))
