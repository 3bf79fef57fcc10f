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
gdev = m._state["entity"]["g"][0].cpu().numpy().copy()
loss.backward()
w1 = m.state_dict()["entity_embedding.weight"].cpu().numpy()
want = g["step1/entity_embedding.weight"]
gref = g["grad1/entity_embedding.weight"]
diff = np.abs(w1 - want)
bad = np.argwhere(diff > 5e-7 + 1e-5 * np.abs(want))
print("loss", float(loss), g["losses"][0])
for r, c in bad:
    roles = [k for k in BATCH_KEYS if k != "relation_id" and k != "user_id" and (g[f"batch1/{k}"] == r).any()]
    print("row", r, "col", c, "diff", diff[r, c], "g ours", gdev[r, c], "g ref", gref[r, c], "w0", g["init/entity_embedding.weight"][r, c], "w1 ours", w1[r, c], "ref", want[r, c], roles)
