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

    def __init__(self, rank, d=6):
        self.n_users, self.n_entities, self.n_relations, self.embedding_size = 11, 17, 4, d
        self._step = 3
        self._touch_bounds = (5, 40, 3)   # entity bound above the table size: capacity clamps to the rows
        rng = np.random.default_rng(100 + rank)
        self.g = []
        self.touched = []
        for rows, parts, n_touch in ((11, 1, 5), (17, 2, 9), (4, 1, 2)):
            g = np.zeros((rows, parts * d), dtype=np.float32)
            ids = np.sort(rng.choice(rows, size=n_touch, replace=False))
            g[ids] = rng.standard_normal((n_touch, parts * d)).astype(np.float32)
            self.g.append(torch.from_numpy(g))
            self.touched.append(torch.from_numpy(ids))

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
        assert ex.layout.caps == [5, 17, 3]
        assert ex.bytes_per_step == ex.layout.nbytes
        np.save(os.path.join(out_dir, f"before_{rank}.npy"), np.concatenate([b.numpy().ravel() for b in before]))
        np.save(os.path.join(out_dir, f"after_{rank}.npy"), np.concatenate([g.numpy().ravel() for g in model.g]))
        # a second step with another batch shape re-plans the buffers
        model2 = _FakeModel(rank + 10)
        model2._touch_bounds = (7, 12, 2)
        ex(model2)
        assert ex.layout.caps == [7, 12, 2]
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
