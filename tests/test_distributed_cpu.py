"""World-size-2 `gloo` tests of the data-parallel host logic (no GPU): the row-sparse exchange
plumbing with CPU stand-ins for the pack / add kernels, user-block sharding and the exact
metric-sum reduction.  The CUDA pack / add kernels themselves are covered by the -m gpu tests."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hopwise_b200.distributed import FlatLayout, RowSparseExchange, reduce_metric_sums, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeModel:
    """What RowSparseExchange reads from a FusedKGEModel, with dense CPU gradient accumulators."""

    USER_TABLES = ("u",)
    ENTITY_TABLES = ("e_re", "e_im")
    RELATION_TABLES = ("r",)

    def __init__(self, rank, d=6, rows=(11, 17, 4)):
        self.n_users, self.n_entities, self.n_relations, self.embedding_size = *rows, d
        self._step = 3
        self._touch_bounds = (5, 40, 3)   # entity bound above the table size: capacity clamps to the rows
        rng = np.random.default_rng(100 + rank)
        self.g = []
        self.touched = []
        for nrow, parts, n_touch in ((rows[0], 1, 5), (rows[1], 2, 9), (rows[2], 1, 2)):
            g = np.zeros((nrow, parts * d), dtype=np.float32)
            ids = np.sort(rng.choice(nrow, size=n_touch, replace=False))
            g[ids] = rng.standard_normal((n_touch, parts * d)).astype(np.float32)
            self.g.append(torch.from_numpy(g))
            self.touched.append(torch.from_numpy(ids))
        # the flat buffers of the dense route (recommender.py::_ensure_state)
        self._state = {"g_flat": torch.cat([g.reshape(-1) for g in self.g]),
                       "row_state_flat": torch.full((sum(rows), 2), -1, dtype=torch.int32)}
        g_off = rs_off = 0
        for fam, g, ids in zip(("user", "entity", "relation"), self.g, self.touched):
            self._state[fam] = {"g_span": (g_off, g_off + g.numel()), "rs_span": (rs_off, rs_off + g.shape[0])}
            self._state["row_state_flat"][rs_off + ids, 1] = self._step + 1
            g_off += g.numel()
            rs_off += g.shape[0]
        self.g = [self._state["g_flat"][self._state[f]["g_span"][0]:self._state[f]["g_span"][1]].view_as(g)
                  for f, g in zip(("user", "entity", "relation"), self.g)]

    def parameters(self):
        return iter([torch.zeros(1)])


def _cpu_pack(model, which, step, count, ids, rows):
    t = model.touched[which]
    n = t.numel()
    count[0] = n
    ids[:n] = t
    rows[:n] = model.g[which][t]
    model.g[which][t] = 0.0          # the kernel hands the rows over and zeroes them
    model.touched[which] = t[:0]


def _cpu_add(model, which, step, count, ids, rows):
    n = int(count[0])
    model.g[which].index_add_(0, ids[:n], rows[:n])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _FakeModel(rank)
        before = [g.clone() for g in model.g]
        ex = RowSparseExchange(model, pack_fn=_cpu_pack, add_fn=_cpu_add, device=torch.device("cpu"))
        ex(model)
        # user: 5 of 11 rows can be touched -> sparse lists; entity (bound above the table) and relation -> dense
        assert ex.dense == [False, True, True] and ex.layout.caps == [5, 1, 1]
        assert ex.bytes_per_step == ex.layout.nbytes + (17 * 2 * 6 + 4 * 6) * 4
        marks = model._state["row_state_flat"][11:, 1]
        np.save(os.path.join(out_dir, f"marks_{rank}.npy"), marks.numpy())
        np.save(os.path.join(out_dir, f"before_{rank}.npy"), np.concatenate([b.numpy().ravel() for b in before]))
        np.save(os.path.join(out_dir, f"after_{rank}.npy"), np.concatenate([g.numpy().ravel() for g in model.g]))
        # the bounds the ranks agreed on at the first step stay in force: a local batch that can touch more rows than
        # agreed is refused (ranks planning with different shapes would issue mismatched collectives) ...
        model2 = _FakeModel(rank + 10, rows=(44, 68, 16))
        model2._touch_bounds = (6, 9, 2)
        try:
            ex(model2)
            raise AssertionError("a batch beyond the agreed bounds must be refused")
        except RuntimeError as exc:
            assert "max_batch_rows" in str(exc)
        # ... a smaller one plans with the agreed bounds, and after reset_bounds() the ranks agree again: rank 1's
        # batch can touch more user rows than rank 0's, both plan with the maximum
        model2._touch_bounds = (5, 9, 2)
        ex(model2)
        assert ex.dense == [False, True, False] and ex.layout.caps == [5, 1, 3]
        model3 = _FakeModel(rank + 20, rows=(44, 68, 16))
        model3._touch_bounds = (4 + rank, 9, 2)
        ex.reset_bounds()
        ex(model3)
        assert ex.dense == [False, False, False] and ex.layout.caps == [5, 9, 2]
        # no multicast memory on a CPU fabric: the probe agrees on False, and "auto" keeps the plain exchange
        from hopwise_b200.distributed import multicast_available

        assert multicast_available(torch.device("cpu")) is False
        # exact metric means from per-rank sums
        sums = torch.tensor([[1.0 + rank, 2.0], [3.0, 4.0 * (rank + 1)]], dtype=torch.float64)
        tot, n = reduce_metric_sums(sums, n_users=10 + rank)
        assert n == 21
        np.testing.assert_allclose(tot.numpy(), [[3.0, 4.0], [6.0, 12.0]])
    finally:
        dist.destroy_process_group()


def test_row_sparse_exchange_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    before = [np.load(tmp_path / f"before_{r}.npy") for r in range(world)]
    after = [np.load(tmp_path / f"after_{r}.npy") for r in range(world)]
    # every rank ends with the rank-ordered sum of all ranks' gradients, bit-identical across ranks
    want = before[0] + before[1]
    np.testing.assert_array_equal(after[0], after[1])
    np.testing.assert_array_equal(after[0], want)
    # dense route: every row of a dense table is marked (a row nobody touched takes a zero-gradient Adam step)
    m0, m1 = np.load(tmp_path / "marks_0.npy"), np.load(tmp_path / "marks_1.npy")
    np.testing.assert_array_equal(m0, m1)
    assert (m0 == 4).all()


def test_flat_layout_views_do_not_overlap():
    lay = FlatLayout([5, 9, 2], [1, 2, 1], 10)
    buf = torch.zeros(lay.nbytes, dtype=torch.uint8)
    for which in range(3):
        count, ids, rows = lay.views(buf, which)
        count.fill_(which + 1)
        ids.fill_(100 + which)
        rows.fill_(float(which) + 0.5)
    for which in range(3):
        count, ids, rows = lay.views(buf, which)
        assert int(count[0]) == which + 1 and bool((ids == 100 + which).all()) and bool((rows == which + 0.5).all())
        assert rows.shape == (lay.caps[which], lay.parts[which] * 10)
    assert lay.nbytes % 16 == 0
    assert all(rows % 16 == 0 and ids % 8 == 0 for _, ids, rows in lay.offsets)   # float4 / int64 access on the GPU


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 1_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
