"""Split the ncu source-page samples of a warp-specialised kernel into regions delimited by USETMAXREG (role entry
points) and print per-region stall mix + top instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 12
sections, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; sections.append(cur)
    elif r and r[0] == "Address":
        hdr = r; cur["hdr"] = r
    elif cur is not None and hdr is not None and len(r) == len(hdr):
        cur["rows"].append(r)
sec = sections[which]
hdr = sec["hdr"]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions, cur = [], {"name": "prologue", "rows": []}
for r in sec["rows"]:
    if "USETMAXREG" in r[ix["Source"]]:
        regions.append(cur); cur = {"name": r[ix["Source"]].strip(), "rows": []}
    cur["rows"].append(r)
regions.append(cur)
tot = sum(int(r[ix["# Samples"]]) for r in sec["rows"])
for reg in regions:
    n = sum(int(r[ix["# Samples"]]) for r in reg["rows"])
    ex = sum(int(r[ix["Instructions Executed"]]) for r in reg["rows"])
    agg = {h: sum(int(r[ix[h]]) for r in reg["rows"]) for h in stalls}
    print("=== region %-45s samples %6d (%4.1f%%)  warp-instr executed %d" % (reg["name"][:45], n, 100 * n / tot, ex))
    print("    stalls:", ", ".join("%s %.0f%%" % (h[6:], 100 * v / max(n, 1)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    for r in sorted(reg["rows"], key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
        s = int(r[ix["# Samples"]])
        st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print("    %6d  %s  %-60s %s" % (s, r[ix["Address"]][-5:], r[ix["Source"]].strip()[:60], st))
