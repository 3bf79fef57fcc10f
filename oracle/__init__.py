"""CPU oracle for the KGE hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithms that tail-unica/hopwise runs for
TransE / RotatE / DistMult / ComplEx (and TorusE / TransH / TransD) training, KG negative sampling and full-sort
top-k evaluation.  It is the checker the CUDA path is compared against; it is never
the thing that is shipped or measured as the product.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``hopwise_b200/`` imports it,
and ``tests/test_abi_cpu.py::test_product_has_no_cpu_fallback_and_no_oracle_import`` enforces that.

Parity status: PINNED.  Every function here is checked against outputs of the
reference itself (hopwise v0.9.1.post1 imported from /root/reference in the build
container; fixtures under ``tests/golden/`` made by ``tests/golden/make_golden.py``)
and against the reference's own metric known-answer vectors
(/root/reference/tests/metrics/test_topk_metrics.py:26-108).

Third-party arithmetic the reference delegates to (not under /root/reference):
  * torch (uv.lock pins 2.7.1; container has 2.11.0): nn.Embedding, TripletMarginLoss,
    MarginRankingLoss, BCEWithLogitsLoss, optim.Adam, topk.
  * numpy (uv.lock pins 2.1.3; container has 2.3.5): legacy global MT19937
    ``np.random.seed`` / ``np.random.randint`` (masked rejection sampling).
"""
