"""Dev tool: split the SASS of each kernel in an `ncu --page source --csv` export at WARPSYNC / BAR instructions and
print, per segment, the share of stall samples and executed instructions plus the dominant stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:60], "hdr": None, "rows": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for s in secs:
    h = s["hdr"]
    isamp, isrc, iinst = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp]) for r in s["rows"]) or 1
    toti = sum(int(r[iinst]) for r in s["rows"]) or 1
    print("==", s["name"], "samples", tot, "warp-instructions", toti)
    seg = {"samp": 0, "inst": 0, "n": 0, "stalls": {}, "ops": {}}
    k = 0

    def flush(tag):
        global seg, k
        if seg["n"]:
            top = sorted(seg["stalls"].items(), key=lambda kv: -kv[1])[:3]
            ops = sorted(seg["ops"].items(), key=lambda kv: -kv[1])[:5]
            print(f"  seg{k:2d} {seg['n']:4d} sass  samples {100 * seg['samp'] / tot:5.1f}%  inst {100 * seg['inst'] / toti:5.1f}%  "
                  f"{[(a.replace('stall_', ''), round(100 * b / max(seg['samp'], 1))) for a, b in top]}  {ops}  -> {tag}")
        seg = {"samp": 0, "inst": 0, "n": 0, "stalls": {}, "ops": {}}
        k += 1

    for r in s["rows"]:
        op = r[isrc].split()
        op = [t for t in op if not t.startswith("@")]
        name = op[0].split(".")[0] if op else "?"
        seg["samp"] += int(r[isamp])
        seg["inst"] += int(r[iinst])
        seg["n"] += 1
        seg["ops"][name] = seg["ops"].get(name, 0) + int(r[iinst])
        for i, c in stall_cols:
            v = int(r[i]) if r[i].isdigit() else 0
            if v:
                seg["stalls"][c] = seg["stalls"].get(c, 0) + v
        if name in ("WARPSYNC", "BAR", "BRA", "EXIT") and (name != "BRA" or seg["n"] > 40):
            flush(r[isrc].strip()[:40])
    flush("end")
