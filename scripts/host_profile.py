"""Where the host time of one fused step goes (cProfile over 3000 small steps; GPU work is negligible at B=2048)."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
w = bench.WORKLOADS["cfg2_transe_ml1m_b2048"]
dev = torch.device("cuda", 0)
model = bench.make_model(w, dev)
batches = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in bench.synth_batches(w, 4, 1)]
def run(n, sync, one_call):
    for i in range(n):
        if one_call:
            loss = model.train_step(batches[i % 4])
            if sync:
                loss.item()
        else:
            loss = model.calculate_loss(batches[i % 4])
            if sync:
                loss.item()
            loss.backward()
run(200, True, False)
run(200, True, True)
torch.cuda.synchronize()
for one_call in (False, True):
    for sync in (False, True):
        t0 = time.perf_counter(); run(3000, sync, one_call); torch.cuda.synchronize()
        print(f"one_call={one_call} sync={sync}: {(time.perf_counter() - t0) / 3000 * 1e6:.1f} us/step")
pr = cProfile.Profile(); pr.enable(); run(3000, False, True); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
