import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from hopwise_b200.loader import pack_batch
w = bench.WORKLOADS["cfg2_transe_ml1m"]
dev = torch.device("cuda", 0)
model = bench.make_model(w, dev)
host = [{k: torch.from_numpy(v).pin_memory() for k, v in b.items()} for b in bench.synth_batches(w, 4, 1)]
packed = [pack_batch(b) for b in host]
side = torch.cuda.Stream(dev)
bufs = [torch.empty_like(packed[0].base, device=dev) for _ in range(3)]
N = 40
def loop(label, copy, compute, sync):
    torch.cuda.synchronize()
    ts = []
    evs = [None] * 3
    t_all = time.perf_counter()
    for i in range(N):
        t0 = time.perf_counter()
        slot = i % 3
        if copy:
            with torch.cuda.stream(side):
                bufs[slot].copy_(packed[i % 4].base, non_blocking=True)
                e = torch.cuda.Event(); e.record(side); evs[slot] = e
        if compute:
            j = (i - 1) % 3 if copy else 0
            if copy and evs[j] is not None:
                torch.cuda.current_stream().wait_event(evs[j])
            db = packed[0].views(bufs[j])
            l = model.calculate_loss(db)
            if sync: l.item()
            l.backward()
        ts.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t_all) / N * 1e3
    print(f"{label:40s} {tot:.3f} ms/step   host per-iter: median {np.median(ts):.3f} max {np.max(ts):.3f}")
bufs[0].copy_(packed[0].base); bufs[1].copy_(packed[1].base); bufs[2].copy_(packed[2].base)
for rep in range(2):
    loop("copy only (side stream)", True, False, False)
    loop("compute only, no sync", False, True, False)
    loop("compute only, sync", False, True, True)
    loop("copy(side) + compute, no sync", True, True, False)
    loop("copy(side) + compute, sync", True, True, True)
