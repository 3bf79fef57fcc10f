import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_train as T
from conftest import load_golden
g = load_golden("model_TransE_d20.npz")
for rep in range(3):
    m = T._golden_model("TransE", g)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    b = T._gbatch(g, int(g["schedule"][0]))
    loss = T._trainer_step(m, opt, b)
    sd = m.state_dict()
    for k, v in sd.items():
        got = v.cpu().numpy(); want = g[f"step1/{k}"]; init = g[f"init/{k}"]
        diff = np.abs(got - want)
        idx = np.argwhere(diff > 1e-5 * np.abs(want) + 2e-7)
        print("rep", rep, k, "mismatches", len(idx), "max", diff.max())
        for r, c in idx[:8]:
            print("   row", r, "col", c, "got", got[r, c], "want", want[r, c], "init", init[r, c], "grad", g[f"grad1/{k}"][r, c])
