"""Where the time of one device-loader step goes at the reference batch (2048 + 2048): loader alone, step alone, both;
with and without a per-step host sync."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hopwise_b200.loader import DeviceKGLoader
from hopwise_b200.sampler import KGSampler, MTStream, RecSampler

dev = torch.device("cuda", 0)
wx = bench.WORKLOADS["cfg2_transe_ml1m_b2048"]
rng = np.random.default_rng(2024)
iu, ii = rng.integers(1, wx["U"], wx["inters"]), rng.integers(1, wx["I"], wx["inters"])
kh, kr, kt = (rng.integers(1, wx["E"], wx["triples"]), rng.integers(1, wx["R"] - 1, wx["triples"]), rng.integers(1, wx["E"], wx["triples"]))
mt = MTStream(seed=2024, device=dev)
loader = DeviceKGLoader(iu, ii, kh, kr, kt, RecSampler(iu, ii, wx["U"], wx["I"], stream=mt, device=dev),
                        KGSampler(heads=kh, tails=kt, entity_num=wx["E"], stream=mt, device=dev), batch_size=wx["n_rec"], seed=2024, device=dev)
mx = bench.make_model(wx, dev)
it = iter(loader)
for _ in range(10):
    mx.train_step(next(it))
torch.cuda.synchronize()
N = 100

def timed(label, fn, sync):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(N):
        r = fn()
        if sync:
            torch.cuda.synchronize()
    e1.record(); host = time.perf_counter() - t0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f"{label:34s} sync={sync!s:5s} host {host / N * 1e6:7.1f} us  wall {wall / N * 1e6:7.1f} us  device {e0.elapsed_time(e1) / N * 1e3:7.1f} us")

b0 = next(it)
for sync in (False, True):
    timed("loader only", lambda: next(it), sync)
    timed("train_step only (fixed batch)", lambda: mx.train_step(b0), sync)
    timed("loader + train_step", lambda: mx.train_step(next(it)), sync)
    timed("kg sampler only", lambda: loader.kg_sampler.sample_by_entity_ids(b0["head_id"], 1), sync)
    timed("rec sampler only", lambda: loader.rec_sampler.sample_by_user_ids(b0["user_id"], b0["item_id"], 1), sync)
    timed("gather only", lambda: loader._take(loader.kg_order, (loader.kg_head, loader.kg_rel, loader.kg_tail)), sync)

# the bench's sequence: a no-sync loop that keeps every step's loss, then one read at the end
for hold in (False, True, False, True):
    it = iter(loader)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    held = []
    for _ in range(N):
        loss = mx.train_step(next(it))
        if hold:
            held.append(loss)
    host = time.perf_counter() - t0
    if hold:
        total = float(torch.stack(held).double().sum().item())
    torch.cuda.synchronize()
    print(f"no-sync loop, keep losses={hold}: host {host / N * 1e6:.1f} us  wall {(time.perf_counter() - t0) / N * 1e6:.1f} us")
big = torch.empty(2 * 1024 ** 3, dtype=torch.uint8, device=dev)
del big
torch.cuda.empty_cache()
for hold in (False, True):
    it = iter(loader)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    held = []
    for _ in range(N):
        loss = mx.train_step(next(it))
        if hold:
            held.append(loss)
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"after empty_cache, no-sync loop, keep losses={hold}: host {host / N * 1e6:.1f} us  wall {(time.perf_counter() - t0) / N * 1e6:.1f} us")
