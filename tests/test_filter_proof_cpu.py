"""The argument behind the tensor-core full-sort filter (csrc/mma_topk.cu), checked in numpy.

The CUDA path reports exact fp32 scores; the bf16 tensor-core sweep only decides which targets get re-scored.
This file restates that decision rule on the CPU -- bf16-rounded operands, chunk / group maxima, the running
threshold thr = tau - 2 eps with tau = k-th largest chunk maximum among chunks without masked targets, list
compaction when a 128-entry list fills -- and checks on random, heavy-tailed and near-tie data that
  (1) |a^ - a| <= eps_i = 1.02 * 2^-8 * |q_i| * max|t| for every pair, and
  (2) every member of the exact top-k (unmasked targets, score desc / id asc) sits in a group the rule keeps,
      or the row is flagged for the exact path.
It does not run the CUDA code (the -m gpu tests compare that with the fp32 kernel bit for bit); it guards the
rule itself against a change that would make the filter lossy.
"""

import numpy as np
import pytest
import torch

CH, GRP, CAND, ROOM = 32, 4, 128, 16


def _bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def _filter_row(ahat, eps, k, masked, trig=CAND - 5):
    """One row of the sweep: returns (set of kept (chunk, group) pairs, flagged)."""
    n = len(ahat)
    n_chunks = (n + CH - 1) // CH
    pad = np.full(n_chunks * CH, 0.0, dtype=np.float32)   # the image pads the last tile with zero rows
    pad[:n] = ahat
    unsafe = np.zeros(n_chunks, dtype=bool)
    for j in np.nonzero(masked)[0]:
        unsafe[j // CH] = True
    if n % CH:
        unsafe[n // CH:] = True
    entries, thr = [], -np.inf       # entry = (chunk max, chunk id, group mask, unsafe)
    flagged = False

    def compact(entries):
        safe = sorted((e[0] for e in entries if not e[3]), reverse=True)
        tau = safe[k - 1] if len(safe) >= k else -np.inf
        new_thr = np.float32(tau) - np.float32(2.0) * np.float32(eps)
        return [e for e in entries if e[0] >= new_thr], new_thr

    for c in range(n_chunks):
        blk = pad[c * CH:(c + 1) * CH].reshape(CH // GRP, GRP)
        gm = blk.max(axis=1)
        tm = gm.max()
        if tm >= thr:
            entries.append((tm, c, gm >= thr, bool(unsafe[c])))
        if (c % 4 == 3 or c == n_chunks - 1) and len(entries) > trig:   # checked once per 128-target tile
            entries, thr = compact(entries)
            if len(entries) > CAND - ROOM:
                return set(), True
    # re-score side: final tau over everything that survived, then the groups of the surviving chunks
    safe = sorted((e[0] for e in entries if not e[3] and e[0] >= thr), reverse=True)
    tau = safe[k - 1] if len(safe) >= k else -np.inf
    final_thr = max(thr, np.float32(tau) - np.float32(2.0) * np.float32(eps))
    kept = {(e[1], g) for e in entries if e[0] >= final_thr for g in np.nonzero(e[2])[0]}
    return kept, flagged


# enough targets for several compactions per row (a list fills after 123 appended chunks of 32 targets)
CASES = [("normal", 64, 20000), ("heavy_tail", 64, 20000), ("near_ties", 32, 12000), ("dist", 100, 16000),
         ("tiny", 16, 700)]


@pytest.mark.parametrize("kind,d,n", CASES)
def test_filter_keeps_the_exact_topk(kind, d, n):
    rng = np.random.default_rng(sum(map(ord, kind)))   # (hash() is salted per process)
    rows, k = 12, 20
    T = rng.standard_normal((n, d)).astype(np.float32) * np.float32(0.1)
    Q = rng.standard_normal((rows, d)).astype(np.float32) * np.float32(0.1)
    if kind == "heavy_tail":
        T[rng.integers(0, n, 30)] *= 3.0     # (x25 makes eps swallow the score spread: every row is flagged)
        T[rng.integers(0, n, 10)] *= -4.0
    if kind == "near_ties":          # 64 targets within a few bf16 ulps of each other at the top, spread over chunks
        base = rng.standard_normal(d).astype(np.float32)
        idx = rng.choice(n, 64, replace=False)
        T[idx] = base + rng.standard_normal((64, d)).astype(np.float32) * np.float32(1e-4)
        Q[:] = base * np.float32(0.5) + Q * np.float32(0.01)
    dist = kind == "dist"
    a = Q.astype(np.float64) @ T.astype(np.float64).T
    ahat = (_bf16(Q).astype(np.float32) @ _bf16(T).astype(np.float32).T).astype(np.float32)
    tn2 = (T.astype(np.float64) ** 2).sum(1)
    tmax = np.sqrt(tn2.max()) * 1.0001
    if dist:                          # -|q - t|^2 = 2 (q.t - |t|^2 / 2) - |q|^2: the sweep ranks q.t - |t|^2 / 2
        a = a - 0.5 * tn2[None, :]
        s = (-0.5 * tn2).astype(np.float32)
        hi = _bf16(s)
        mid = _bf16(s - hi)
        lo = _bf16(s - hi - mid)
        ahat = (ahat + (hi + mid + lo)[None, :]).astype(np.float32)
    for i in range(rows):
        nq = np.sqrt((Q[i].astype(np.float64) ** 2).sum())
        eps = 1.02 * 2.0 ** -8 * nq * tmax + (2.0 ** -20 * 0.5 * tmax * tmax if dist else 0.0)
        assert np.abs(ahat[i].astype(np.float64) - a[i]).max() <= eps, "error bound violated"
        masked = np.zeros(n, dtype=bool)
        masked[0] = True                                     # the [PAD] target
        masked[rng.integers(1, n, int(rng.integers(0, 60)))] = True   # history
        kept, flagged = _filter_row(ahat[i], eps, k, masked)
        if flagged:
            continue                                         # the row goes to the exact fp32 kernel
        valid = np.nonzero(~masked)[0]
        order = valid[np.lexsort((valid, -a[i, valid]))][:k]  # exact top-k: score desc, id asc
        for j in order:
            assert (j // CH, (j % CH) // GRP) in kept, f"row {i}: exact top-k member {j} was filtered out"


def test_rows_with_fewer_than_k_valid_targets_never_get_a_threshold():
    """With fewer than k chunks free of masked targets tau stays -inf: nothing is ever dropped."""
    rng = np.random.default_rng(3)
    n, k = 640, 20
    ahat = rng.standard_normal(n).astype(np.float32)
    masked = np.ones(n, dtype=bool)
    masked[rng.integers(1, n, 12)] = False                  # 12 valid targets only
    kept, flagged = _filter_row(ahat, 1e-3, k, masked)
    assert flagged or all((j // CH, (j % CH) // GRP) in kept for j in np.nonzero(~masked)[0])
