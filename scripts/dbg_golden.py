import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_train as T
from conftest import load_golden
g = load_golden("model_TransE_d20.npz")
for rep in range(3):
    m = T._golden_model("TransE", g)
    b = T._gbatch(g, int(g["schedule"][0]))
    m._launch_forward(b, with_grad=True)
    torch.cuda.synchronize()
    gu = m._state["user"]["g"][0].cpu().numpy()
    want = g["grad1/user_embedding.weight"]
    diff = np.abs(gu - want)
    idx = np.argwhere(diff > 1e-4 * np.abs(want) + 1e-8)
    print("rep", rep, "grad mismatches", len(idx), "max abs", diff.max(), "max |want|", np.abs(want).max())
    for r, c in idx[:6]:
        print("   row", r, "col", c, "got", gu[r, c], "want", want[r, c])
    users = b["user_id"].cpu().numpy()
    if len(idx):
        r = idx[0][0]
        print("   user", r, "appears", (users == r).sum(), "times in the rec half")
    m.flush()
