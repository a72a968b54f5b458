"""Exact aggregates at the BASELINE.json sizes -- 10 M (configs 0-1), 100 M (config 4) and 1 B rows (configs 2-3) -- against
the oracle, not only against size-independent properties.

The synthetic table is generated on the device (k_synth) and, independently, on the host cores by the oracle's twin of the
generator (bit-identical rows); the checkers are
  * the reference's own arithmetic: one double, strictly left to right in id order (custom_bplus_db.cpp:242-251 sum_amount,
    :263-274 sum_amount_where) -- stated tolerance relative <= 1e-12;
  * an effectively exact sum (long double Neumaier accumulation) -- stated tolerance <= 4 ulp;
  * COUNT exact.
The 1 B-row case holds the 8 GB column in host memory for a few seconds."""
import math

import pytest

import approximatequeryengine_b200 as aqe

pytestmark = pytest.mark.gpu

LO, HI = 100.0, 500.0


@pytest.mark.parametrize("n", [10_000_000, 100_000_000, 1_000_000_000])
def test_exact_aggregates_against_the_oracle_at_config_sizes(oracle, n, monkeypatch):
    monkeypatch.delenv("AQE_MIN_SHARD_ROWS", raising=False)
    e = aqe.Engine(0).generate(n, seed=7, columns=("amount",))
    col = oracle.synth_amount(n, seed=7)
    # spot-check that the two generators agree (the 200 k-row byte comparison lives in test_device_generator_bit_identical)
    for first in (0, n // 2 - 50, n - 100):
        assert (e.read_column("amount", first, 100) == col[first:first + 100]).all()
    serial, cnt = oracle.sum_col_serial(col)
    exact, _ = oracle.sum_col_exact(col)
    s = e.sum_amount()
    assert cnt == n == e.count
    assert abs(s - serial) <= 1e-12 * abs(serial), (n, s, serial)          # the reference's serial sum, stated tolerance
    assert abs(s - exact) <= 4 * math.ulp(exact), (n, s, exact)            # the exactly rounded sum
    assert e.sum_amount() == s                                              # bit stable
    serial_w, cnt_w = oracle.sum_col_serial(col, (LO, HI))
    exact_w, cnt_e = oracle.sum_col_exact(col, (LO, HI))
    w, c = e.sum_amount_where(LO, HI)
    assert c == cnt_w == cnt_e, (n, c, cnt_w)                               # COUNT with the range predicate: exact
    assert abs(w - serial_w) <= 1e-12 * abs(serial_w) and abs(w - exact_w) <= 4 * math.ulp(exact_w), (n, w, serial_w, exact_w)
    p = e.scan("amount", "amount", LO, HI)
    assert (p.count, p.sum) == (c, w)
    # the multithreaded restatement used as the CPU baseline agrees too (region sums in double, ordered merge)
    mt, cmt = oracle.scan_mt(col, aos=False, threads=16, pred=(LO, HI))
    assert cmt == c and abs(mt - w) <= 1e-12 * abs(w)
    # the same table as a sharded handle (colocated shards on this GPU; peer mode with more GPUs): same count, <= 1 ulp
    devs = [0, 0] if aqe_gpus() < 2 else list(range(min(aqe_gpus(), 8)))
    del e
    g = aqe.Engine(devices=devs).generate(n, seed=7, columns=("amount",))
    assert g.shard_count == (len(devs) if n // len(devs) >= 1 << 24 else max(1, n >> 24))
    wg, cg = g.sum_amount_where(LO, HI)
    assert cg == c and abs(wg - w) <= math.ulp(w) and abs(g.sum_amount() - s) <= math.ulp(s)
    g.close()


def aqe_gpus():
    c = aqe.C.c_int()
    aqe.lib().aqe_device_count(aqe.C.byref(c))
    return c.value
