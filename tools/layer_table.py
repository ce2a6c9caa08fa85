"""Renders gpurun_out/layer_table_<config>.json (written by bench.py's roofline leg) as the per-layer markdown table kept under
profiles/.      python tools/layer_table.py gpurun_out/layer_table_fcgan.json > profiles/r2_layer_table_fcgan.md"""
import json
import sys

d = json.load(open(sys.argv[1]))
rows = d["rows"]
peak, hbm = d.get("peak_tflops"), d.get("hbm_gbs")
tot_c = sum(r["cold_us"] * r["calls_per_step"] for r in rows)
tot_w = sum(r["warm_us"] * r["calls_per_step"] for r in rows)
print("# Per-call table of one step (%s, batch %d): graph step %.3f ms; instrumented calls sum to %.3f ms cold / %.3f ms warm"
      % (d["precision"], d["batch"], d["step_ms"], tot_c * 1e-3, tot_w * 1e-3))
print()
print("Every distinct kernel call of one step replayed 10x from its own CUDA graph (CUDA events on the launching stream): cold = a")
print("256 MB buffer rewritten before each launch (rewrite time subtracted), warm = back to back.  tf32 peak %.0f TFLOP/s (measured"
      % (peak or 0))
print("live, torch.matmul), HBM %.0f GB/s (MEASURED_PEAKS.json).  GB/s is against ALGORITHMIC bytes." % (hbm or 0))
print()
print("| call | kernel | n/step | cold us | warm us | us/step (cold) | TFLOP/s cold (warm) | % tf32 peak | GB/s alg. | % HBM |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in rows:
    tf = "%.0f (%.0f)" % (r["tflops_cold"], r["tflops_warm"]) if r.get("tflops_cold") else ""
    pf = "%.0f" % (100 * r["tflops_cold"] / peak) if r.get("tflops_cold") and peak else ""
    gb = "%.0f" % r["gbs_cold"] if r.get("gbs_cold") else ""
    pg = "%.0f" % (100 * r["gbs_cold"] / hbm) if r.get("gbs_cold") and hbm else ""
    print("| %s | %s | %d | %.1f | %.1f | %.1f | %s | %s | %s | %s |" % (r["call"], r["kernel"], r["calls_per_step"], r["cold_us"],
          r["warm_us"], r["cold_us"] * r["calls_per_step"], tf, pf, gb, pg))
