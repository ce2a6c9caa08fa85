"""Markdown per-call table from gpurun_out/layer_table_<config>.json (written by bench.py's roofline leg).
    python tools/layer_table.py gpurun_out/layer_table_fcgan.json > profiles/r2_layer_table_fcgan.md"""
import json, sys
t = json.load(open(sys.argv[1]))
rows, peak, hbm = t["rows"], t["peak_tflops"], t["hbm_gbs"]
tot = sum(r["us"] * r["calls_per_step"] for r in rows)
print("# Per-call table of one step (%s, batch %d): graph step %.3f ms; instrumented calls sum to %.3f ms\n"
      % (t["precision"], t["batch"], t["step_ms"], tot * 1e-3))
print("Every distinct kernel call of one step replayed 10x back to back from its own CUDA graph (CUDA events on the launching stream,")
print("best of 3). Calls whose tensors fit the 126 MB L2 run L2-warm, as they do inside the step right after their producer.")
print("tf32 peak %.0f TFLOP/s (measured live), HBM %.0f GB/s (MEASURED_PEAKS.json).\n" % (peak, hbm))
print("| call | kernel | n/step | us | us/step | TFLOP/s | % tf32 peak | GB/s alg. | % HBM | alg. MB |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in rows:
    tf = "%.0f" % r["tflops"] if r.get("tflops") else ""
    pf = "%.0f" % (100 * r["tflops"] / peak) if r.get("tflops") and peak else ""
    gb = "%.0f" % r["gbs"] if r.get("gbs") else ""
    pg = "%.0f" % (100 * r["gbs"] / hbm) if r.get("gbs") and hbm else ""
    print("| %s | %s | %d | %.1f | %.1f | %s | %s | %s | %s | %.0f |" % (r["call"], r["kernel"], r["calls_per_step"], r["us"],
          r["us"] * r["calls_per_step"], tf, pf, gb, pg, r["alg_mbytes"]))
