"""GPU time per step of the fused step at the reference batch as the step count grows (lazy-Adam replay cost)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_transe_ml1m_b2048"
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
model = bench.make_model(w, dev)
batches = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in bench.synth_batches(w, 8, 1)]
done = 0
for chunk in (50, 200, 500, 1000, 2000):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(chunk):
        loss = model.calculate_loss(batches[i % 8])
        loss.backward()
    e1.record()
    torch.cuda.synchronize()
    done += chunk
    print(f"steps {done - chunk:5d}..{done:5d}: gpu {e0.elapsed_time(e1) / chunk * 1e3:7.1f} us/step, wall {(time.perf_counter() - t0) / chunk * 1e6:7.1f} us/step")
