import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from kge_helpers import BATCH_KEYS, make_product_model
g = load_golden("model_TransH_d20.npz")
U, I, E, R, d = (int(x) for x in g["shape"])
m = make_product_model("TransH", U, I, E, R, d, margin=float(g["margin"]))
m.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init/")}, strict=True)
b = {k: torch.from_numpy(g[f"batch1/{k}"]).cuda() for k in BATCH_KEYS}
loss = m.calculate_loss(b)
gdevs = {"user_embedding": m._state["user"]["g"][0].cpu().numpy().copy(), "entity_embedding": m._state["entity"]["g"][0].cpu().numpy().copy(),
         "relation_embedding": m._state["relation"]["g"][0].cpu().numpy().copy(), "norm_vec": m._state["relation"]["g"][1].cpu().numpy().copy()}
loss.backward()
print("loss", float(loss), g["losses"][0])
sd = m.state_dict()
for t, gdev in gdevs.items():
    w1 = sd[t + ".weight"].cpu().numpy()
    want = g[f"step1/{t}.weight"]
    gref = g[f"grad1/{t}.weight"]
    diff = np.abs(w1 - want)
    bad = np.argwhere(diff > 5e-7 + 1e-5 * np.abs(want))
    print(t, "bad", len(bad), "max |g ours - g ref|", np.abs(gdev - gref).max())
    for r, c in bad:
        roles = [k for k in BATCH_KEYS if (g[f"batch1/{k}"] == r).any()]
        print("  row", r, "col", c, "diff", diff[r, c], "g ours", gdev[r, c], "g ref", gref[r, c], roles)
        if t == "user_embedding":
            idx = np.flatnonzero(g["batch1/user_id"] == r)
            print("   triples:", [(int(g["batch1/item_id"][i]), int(g["batch1/neg_item_id"][i])) for i in idx])
