"""Where the end-to-end step time goes: H2D alone, compute alone (with the per-step loss.item()), both."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from hopwise_b200.loader import DevicePrefetcher

w = bench.WORKLOADS["cfg2_transe_ml1m"]
dev = torch.device("cuda", 0)
model = bench.make_model(w, dev)
host = [{k: torch.from_numpy(v).pin_memory() for k, v in b.items()} for b in bench.synth_batches(w, 4, 1)]
devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
N = 30

def timed(fn, label):
    for _ in range(3): fn(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(N): fn(i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N * 1e3
    print(f"{label:46s} {dt:.3f} ms/step"); return dt

def h2d(i):
    return {k: v.to(dev, non_blocking=True) for k, v in host[i % 4].items()}
def comp_sync(i):
    l = model.calculate_loss(devb[i % 4]); l.item(); l.backward()
def comp_nosync(i):
    l = model.calculate_loss(devb[i % 4]); l.backward()
def serial(i):
    db = h2d(i); l = model.calculate_loss(db); l.item(); l.backward()
timed(h2d, "H2D of 7 id vectors (14.7 MB, pinned)")
timed(comp_nosync, "compute, ids resident, no host sync")
timed(comp_sync, "compute, ids resident, loss.item() per step")
timed(serial, "serial: H2D + compute + item()")
for depth in (1, 2, 3):
    pf = DevicePrefetcher([host[i % 4] for i in range(N)], dev, depth=depth)
    def run():
        for db in pf:
            l = model.calculate_loss(db); l.item(); l.backward()
    run(); torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize()
    print(f"prefetch depth {depth}: {(time.perf_counter()-t0)/N*1e3:.3f} ms/step")
pf2 = DevicePrefetcher([host[i % 4] for i in range(N)], dev, depth=2)
def run2():
    for db in pf2:
        l = model.calculate_loss(db); l.backward()
run2(); torch.cuda.synchronize(); t0 = time.perf_counter(); run2(); torch.cuda.synchronize()
print(f"prefetch depth 2, no item(): {(time.perf_counter()-t0)/N*1e3:.3f} ms/step")
from hopwise_b200.loader import pack_batch
packed = [pack_batch(b) for b in host]
def h2d_packed(i):
    return packed[i % 4].to(dev)
timed(h2d_packed, "H2D of one packed buffer (14.7 MB, pinned)")
pf3 = DevicePrefetcher([packed[i % 4] for i in range(N)], dev, depth=2)
for rep in range(3):
    def run3():
        for db in pf3:
            l = model.calculate_loss(db); l.item(); l.backward()
    torch.cuda.synchronize(); t0 = time.perf_counter(); run3(); torch.cuda.synchronize()
    print(f"packed prefetch depth 2 (rep {rep}): {(time.perf_counter()-t0)/N*1e3:.3f} ms/step")
