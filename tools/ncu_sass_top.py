"""Top SASS lines by stall samples from `ncu --page source --csv --print-source sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print("total samples", tot, "warp instructions", inst, "lines", len(body))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
top = sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]
for r in top:
    st = {h[6:]: int(r[ix[h]] or 0) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%6s %5.1f%% exec %9s shW %9s/%-9s  %-60s %s" % (r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]] or 0) / max(tot, 1),
          r[ix["Instructions Executed"]], r[ix["L1 Wavefronts Shared"]], r[ix["L1 Wavefronts Shared Ideal"]], r[ix["Source"]][:60], st))
# instruction mix
mix = {}
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    mix[op] = mix.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
print("mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(inst, 1)) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14]))
shw = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in body); shi = sum(int(r[ix["L1 Wavefronts Shared Ideal"]] or 0) for r in body)
print("shared wavefronts", shw, "ideal", shi, " global L2 sectors", sum(int(r[ix["L2 Theoretical Sectors Global"]] or 0) for r in body),
      "ideal", sum(int(r[ix["L2 Theoretical Sectors Global Ideal"]] or 0) for r in body))
