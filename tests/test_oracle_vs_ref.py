"""The oracle (oracle/aqe_oracle.c) against the UNMODIFIED reference compiled in place (oracle/_ref/libaqe_ref.so).
Runs only where that library exists (this container: /root/reference is present and `make -C oracle ref` built it;
elsewhere the golden vectors minted from it carry the pin -- tests/test_oracle_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import Ref, RefScheduler, make_params

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libaqe_ref.so not built (no /root/reference here)")

DETERMINISTIC = ["memory_stride", "slow_pointer", "fast_pointer", "dual_pointer", "parallel_pointer", "random_pointer",
                 "optimized_address_arithmetic", "index_based", "byte_offset", "optimized_clt", "block", "page", "parallel_block",
                 "node_skip", "balanced_tree", "direct_access", "adaptive_block", "stratified_block"]


@pytest.mark.parametrize("n", [254, 255, 256, 381, 382, 1000, 12345, 65537, 99999])
def test_index_sets_and_exact_sums_match_reference(oracle, n):
    rows = oracle.synth(n, seed=11)
    R = Ref(rows)
    assert R.total() == n and R.sum_amount() == oracle.sum_amount(rows)          # bit-equal to the serial id-order sum
    assert R.avg_amount() == oracle.avg_amount(rows)
    for lo, hi in ((100.0, 500.0), (0.0, 1.0), (999.0, 1e9), (500.0, 100.0)):
        assert R.sum_amount_where(lo, hi) == oracle.sum_amount_where(rows, lo, hi)[0]
    assert R.tree_height() == oracle.tree_height(n) and R.node_count() == oracle.node_count(n)
    order = None
    for m in DETERMINISTIC:
        if m == "memory_stride" and 255 <= n < 1000:
            continue                           # reference result depends on earlier calls there (stale subtree count)
        for p in (0.1, 1.0, 7.5, 20.0, 50.0, 100.0):
            for kw in ({}, {"num_threads": 3, "step_size": 3, "block_size": 1024 if m == "page" else 100,
                            "block_size_max": 5 if m == "stratified_block" else 900, "seed": 7}):
                if m == "dual_pointer" and int(n * p / 100.0) < 3:
                    continue                   # reference divides by zero
                prm = make_params(m, p, **kw)
                want = R.sample(m, prm)["id"] - 1
                got = oracle.indices(rows, m, prm)
                if m == "stratified_block":
                    if order is None:
                        order = np.argsort(rows["amount"], kind="stable")
                    got = order[got]
                assert len(got) == len(want) and np.array_equal(got, want), (n, m, p, kw)


def test_racy_clt_sampler_is_a_prefix_family_of_the_lockstep_one(oracle):
    """clt_validated_dual_pointer_sample is racy in the reference; every run returns, per thread, a prefix of the same
    stride sequence the lock-step schedule walks, and about as many rows."""
    rows = oracle.synth(200000, seed=5)
    R = Ref(rows)
    prm = make_params("clt_validated_dual_pointer", 20.0, max_error_percent=1.0)
    lock = oracle.indices(rows, "clt_validated_dual_pointer", prm)
    N, T = len(rows), int(len(rows) * 20 / 100.0)
    for _ in range(3):
        got = R.sample("clt_validated_dual_pointer", prm)["id"] - 1
        assert 0.5 * len(lock) <= len(got) <= 2.0 * len(lock) + T
        # each returned row lies on one of the four stride sequences (fast: a + 5k, slow: a + 2 + 5k; regions of N/2)
        half = N // 2
        on_seq = ((got % half) % 5 == 0) | ((got % half) % 5 == 2) | (got % max(1, N // (T // 4)) == 0)
        assert on_seq.all()


def test_scheduler_constants_match_reference(oracle):
    rows = oracle.synth(20000, seed=2)
    S = RefScheduler(rows)
    r = S.run(0)
    assert r.value == oracle.sum_amount(rows) and (r.confidence_level, r.error_margin, r.samples_used, r.status) == (1.0, 0.0, 20000, 0)
    for p, conf in ((10.0, 0.95), (3.0, 0.90), (1.0, 0.85), (0.3, 0.80), (0.1, 0.70)):
        q = S.run(3, "SELECT SUM(amount) FROM sales", p, 4)
        assert (q.confidence_level, q.error_margin, q.samples_used) == (conf, p / 100.0, int(20000 * p / 100.0))
