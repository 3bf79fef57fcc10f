"""One fused full-sort top-k call on the config-4 shape (profiling target; not a benchmark)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kge_helpers import make_product_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="DistMult")
ap.add_argument("--users", type=int, default=8192)
ap.add_argument("--items", type=int, default=200001)
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--path", default="auto")
ap.add_argument("--table-users", type=int, default=100001)
ap.add_argument("--blocks", type=int, default=1)
a = ap.parse_args()
m = make_product_model(a.model, a.table_users, a.items, a.items, 3, a.d)
rng = np.random.default_rng(0)
blocks = []
for _ in range(a.blocks):
    users = torch.from_numpy(rng.integers(1, a.table_users, a.users)).cuda()
    hist = np.sort(rng.integers(1, a.items, (a.users, 50)), axis=1)
    off = torch.arange(0, 50 * a.users + 1, 50, dtype=torch.long).cuda()
    items = torch.from_numpy(hist.reshape(-1)).cuda()
    blocks.append((users, off, items))
for rep in range(a.reps):
    users, off, items = blocks[rep % len(blocks)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids, _ = m.full_sort_topk(users, a.k, off, items, return_scores=False, path=a.path)
    e1.record()
    torch.cuda.synchronize()
    print(f"[{a.path}] fallback_rows={m._mma_last_fallback_rows} ", end="")
    print(f"{a.model} users={a.users} items={a.items} d={a.d} k={a.k}: {e0.elapsed_time(e1):.3f} ms "
          f"-> {a.users / e0.elapsed_time(e1) * 1e3:.0f} users/s")
