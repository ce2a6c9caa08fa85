"""Development: where the end-to-end step time goes (host-side timings of set_input / optimize_parameters / get_current_errors
around the device-resident step), same model as bench.py's default workload."""
import os, sys, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import supervised_gan_b200 as S
S.set_precision("tf32")
args = argparse.Namespace(batch=8, config="fcgan", pool_size=50, no_graph=False, warmup=3)
m = bench.build_model(args, 0, 1)
hb = [(torch.rand(8, 3, 512, 512) * 2 - 1).pin_memory() for _ in range(2)]
m.set_input({"A": hb[0], "A_paths": ["x"]})
for _ in range(8):
    m.optimize_parameters()
torch.cuda.synchronize()
def run(n, sync_each, read):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = [0.0, 0.0, 0.0]
    for i in range(n):
        a = time.perf_counter(); m.set_input({"A": hb[i % 2], "A_paths": ["x"]})
        b = time.perf_counter(); m.optimize_parameters()
        c = time.perf_counter()
        if read: m.get_current_errors()
        elif sync_each: torch.cuda.synchronize()
        d = time.perf_counter(); ts[0] += b - a; ts[1] += c - b; ts[2] += d - c
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    return tot / n * 1e3, [x / n * 1e3 for x in ts]
for name, se, rd in (("replay only, no per-step sync", False, False), ("sync each step", True, False), ("read losses each step (e2e)", False, True)):
    for rep in range(2):
        ms, parts = run(50, se, rd)
        print("%-34s %.3f ms/step | host: set_input %.3f  optimize %.3f  read/sync %.3f" % (name, ms, *parts), flush=True)
