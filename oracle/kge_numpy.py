"""float64 closed forms of the KGE scorers, losses, analytic gradients and Adam
(oracle; test infrastructure only).

These are the formulas the CUDA kernels implement, written out in numpy float64 so a
test can (a) check them against torch autograd on the reference-shaped oracle
(oracle/kge_torch.py) and (b) measure how far fp32 results are from the real-valued
answer when two top-k candidates are nearly tied.

Reference call sites: see oracle/kge_torch.py.  Loss closed forms:
  TransE   mean_i clamp_min(margin + ||a_i - p_i + 1e-6|| - ||a_i - n_i + 1e-6||, 0), a = h + r
           (torch triplet_margin_loss -> pairwise_distance adds eps inside the difference)
  DistMult mean_i clamp_min(margin - s+_i + s-_i, 0)
  RotatE / ComplEx   per task: (sum softplus(-s+) + sum softplus(s-)) / (2 n_task); tasks summed
Adam: torch.optim.Adam single-tensor formulas (betas 0.9/0.999, eps 1e-8, no amsgrad).
"""

from __future__ import annotations

import numpy as np

EPS_PAIRWISE = 1e-6


def _softplus(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def score(model, h, r, t, margin=1.0):
    """h, r, t: lists of float64 arrays (1 or 2 parts) broadcastable on leading dims."""
    if model == "TransE":
        return -np.sqrt(((h[0] + r[0] - t[0]) ** 2).sum(-1))
    if model == "TransD":   # transd.py:86-91, 135-176
        proj = lambda e: e[0] + r[1] * (e[0] * e[1]).sum(-1, keepdims=True)   # noqa: E731
        return -np.sqrt(((proj(h) + r[0] - proj(t)) ** 2).sum(-1))
    if model == "TransH":   # transh.py:53-58, 73-74
        c = 1.0 - r[1].sum(-1, keepdims=True) * r[1]
        return -np.sqrt((((h[0] - t[0]) * c + r[0]) ** 2).sum(-1))
    if model == "TorusE":   # toruse.py:66-76
        frac = lambda a: a - np.trunc(a)   # noqa: E731
        x = (frac(h[0]) + frac(r[0])) - frac(t[0])
        return -(4 * np.minimum(x * x, 1 - x * x).sum(-1))
    if model == "DistMult":
        return (h[0] * r[0] * t[0]).sum(-1)
    if model == "RotatE":
        c, s = np.cos(r[0]), np.sin(r[0])
        re = c * h[0] - s * h[1] - t[0]
        im = c * h[1] + s * h[0] - t[1]
        return margin - np.sqrt((re * re + im * im).sum(-1))
    hr, hi = h
    rr, ri = r
    tr, ti = t
    return (hr * rr * tr + hi * rr * ti + hr * ri * ti - hi * ri * ti).sum(-1)


def pair_loss_and_grads(model, h, r, tp, tn, weight, margin=1.0):
    """Loss contribution and gradients of `weight * pair_loss` for a block of
    (anchor, relation, positive tail, negative tail) rows.

    Returns (loss_sum, gh, gr, gtp, gtn); each g* is a list of arrays shaped like the input.
    """
    if model in ("TransE", "TorusE"):   # (TorusE trains with TransE's objective, toruse.py:81-102)
        x = h[0] + r[0]
        dp = x - tp[0] + EPS_PAIRWISE
        dn = x - tn[0] + EPS_PAIRWISE
        n_p = np.sqrt((dp * dp).sum(-1, keepdims=True))
        n_n = np.sqrt((dn * dn).sum(-1, keepdims=True))
        z = margin + n_p - n_n
        act = (z >= 0).astype(np.float64) * weight
        up = np.where(n_p > 0, dp / np.where(n_p > 0, n_p, 1.0), 0.0)
        un = np.where(n_n > 0, dn / np.where(n_n > 0, n_n, 1.0), 0.0)
        gx = act * (up - un)
        return (np.maximum(z, 0).sum() * weight, [gx], [gx.copy()], [-act * up], [act * un])
    if model == "TransD":
        # transd.py:86-133: TransE on proj(e) = e + r_p * <e, e_p>.  With G the gradient of a projected row:
        # g_e = G + e_p * <G, r_p>,  g_ep = e * <G, r_p>,  g_rp += G * <e, e_p>
        rp = r[1]
        dot = lambda a_, b_: (a_ * b_).sum(-1, keepdims=True)   # noqa: E731
        s_h, s_p, s_n = dot(h[0], h[1]), dot(tp[0], tp[1]), dot(tn[0], tn[1])
        x = h[0] + rp * s_h + r[0]
        dp = x - (tp[0] + rp * s_p) + EPS_PAIRWISE
        dn = x - (tn[0] + rp * s_n) + EPS_PAIRWISE
        n_p = np.sqrt((dp * dp).sum(-1, keepdims=True))
        n_n = np.sqrt((dn * dn).sum(-1, keepdims=True))
        z = margin + n_p - n_n
        act = (z >= 0).astype(np.float64) * weight
        up = np.where(n_p > 0, dp / np.where(n_p > 0, n_p, 1.0), 0.0)
        un = np.where(n_n > 0, dn / np.where(n_n > 0, n_n, 1.0), 0.0)
        gx, gp, gn = act * (up - un), -act * up, act * un
        back = lambda G, e: ([G + e[1] * dot(G, rp), e[0] * dot(G, rp)])   # noqa: E731
        grp = gx * s_h + gp * s_p + gn * s_n
        return (np.maximum(z, 0).sum() * weight, back(gx, h), [gx, grp], back(gp, tp), back(gn, tn))
    if model == "TransH":
        # transh.py:73-107: TransE on rows scaled by c = 1 - sum(w) * w (project(e) = e - (e * w.sum()) * w).
        # d c_j / d w_i = -w_j - sum(w) * [i == j]  =>  g_w = -sum(w) * g_c - <g_c, w>
        w_ = r[1]
        sw = w_.sum(-1, keepdims=True)
        c = 1.0 - sw * w_
        x = h[0] * c + r[0]
        dp = x - tp[0] * c + EPS_PAIRWISE
        dn = x - tn[0] * c + EPS_PAIRWISE
        n_p = np.sqrt((dp * dp).sum(-1, keepdims=True))
        n_n = np.sqrt((dn * dn).sum(-1, keepdims=True))
        z = margin + n_p - n_n
        act = (z >= 0).astype(np.float64) * weight
        up = np.where(n_p > 0, dp / np.where(n_p > 0, n_p, 1.0), 0.0)
        un = np.where(n_n > 0, dn / np.where(n_n > 0, n_n, 1.0), 0.0)
        gx, gtp_, gtn_ = act * (up - un), -act * up, act * un
        gc = gx * h[0] + gtp_ * tp[0] + gtn_ * tn[0]
        gw = -sw * gc - (gc * w_).sum(-1, keepdims=True)
        return (np.maximum(z, 0).sum() * weight, [gx * c], [gx, gw], [gtp_ * c], [gtn_ * c])
    if model == "DistMult":
        sp = (h[0] * r[0] * tp[0]).sum(-1, keepdims=True)
        sn = (h[0] * r[0] * tn[0]).sum(-1, keepdims=True)
        z = margin - sp + sn
        act = (z >= 0).astype(np.float64) * weight
        diff = tn[0] - tp[0]
        hr = h[0] * r[0]
        return (np.maximum(z, 0).sum() * weight, [act * r[0] * diff], [act * h[0] * diff], [-act * hr], [act * hr])
    if model == "RotatE":
        c, s = np.cos(r[0]), np.sin(r[0])
        rot_re = c * h[0] - s * h[1]
        rot_im = c * h[1] + s * h[0]
        loss = 0.0
        gh = [np.zeros_like(h[0]), np.zeros_like(h[1])]
        gr = [np.zeros_like(r[0])]
        gts = []
        for t, sign in ((tp, +1.0), (tn, -1.0)):
            e_re, e_im = rot_re - t[0], rot_im - t[1]
            nrm = np.sqrt((e_re * e_re + e_im * e_im).sum(-1, keepdims=True))
            sc = margin - nrm
            if sign > 0:
                loss += _softplus(-sc).sum() * weight
                dl_ds = -_sigmoid(-sc) * weight
            else:
                loss += _softplus(sc).sum() * weight
                dl_ds = _sigmoid(sc) * weight
            inv = np.where(nrm > 0, 1.0 / np.where(nrm > 0, nrm, 1.0), 0.0)
            q_re, q_im = dl_ds * (-e_re * inv), dl_ds * (-e_im * inv)  # dL/d(residual)
            gh[0] += c * q_re + s * q_im
            gh[1] += -s * q_re + c * q_im
            gr[0] += -q_re * rot_im + q_im * rot_re
            gts.append([-q_re, -q_im])
        return loss, gh, gr, gts[0], gts[1]
    # ComplEx
    hr, hi = h
    rr, ri = r
    loss = 0.0
    gh = [np.zeros_like(hr), np.zeros_like(hi)]
    gr = [np.zeros_like(rr), np.zeros_like(ri)]
    gts = []
    for t, sign in ((tp, +1.0), (tn, -1.0)):
        tr, ti = t
        sc = (hr * rr * tr + hi * rr * ti + hr * ri * ti - hi * ri * ti).sum(-1, keepdims=True)
        if sign > 0:
            loss += _softplus(-sc).sum() * weight
            dl = -_sigmoid(-sc) * weight
        else:
            loss += _softplus(sc).sum() * weight
            dl = _sigmoid(sc) * weight
        gh[0] += dl * (rr * tr + ri * ti)
        gh[1] += dl * (rr * ti - ri * ti)
        gr[0] += dl * (hr * tr + hi * ti)
        gr[1] += dl * (hr * ti - hi * ti)
        gts.append([dl * (hr * rr), dl * (hi * rr + hr * ri - hi * ri)])
    return loss, gh, gr, gts[0], gts[1]


def loss_weights(model, n_rec, n_kg):
    """Per-pair weights of the rec and KG segments in the scalar loss."""
    if model in ("TransE", "DistMult", "TorusE", "TransH", "TransD"):
        w = 1.0 / (n_rec + n_kg)
        return w, w
    return (0.5 / n_rec if n_rec else 0.0), (0.5 / n_kg if n_kg else 0.0)


def adam_dense_step(p, m, v, g, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam, one step on every element (float64 restatement)."""
    m = m + (1 - b1) * (g - m)
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1**step
    bc2 = 1 - b2**step
    p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return p, m, v
