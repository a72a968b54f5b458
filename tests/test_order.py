"""Row order of tables with DUPLICATE ids (include/aqe_b200.h aqe_reference_order, csrc/aqe_order.cpp): where the reference's B+ tree
puts rows that share an id depends on the insert history and on its leaf splits (custom_bplus_db.cpp:32-37, :43-58, :196-231).
Checked against golden orders minted from the unmodified reference (tests/golden/make_order_golden.py) and, where
oracle/_ref/libaqe_ref.so exists, against the live reference on fresh random histories.  Host code only -- no GPU."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import approximatequeryengine_b200 as aqe
from oracle import Ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "order_cases.json")


def test_duplicate_id_order_matches_reference_golden():
    g = json.load(open(GOLDEN))
    for c in g["cases"]:
        got = aqe.reference_order(c["ids"], [tuple(o) for o in c["ops"]])
        assert [int(x) for x in got] == c["order"], c["name"]
    for c in g["big"]:
        ids = np.random.default_rng(c["seed"]).integers(0, c["hi"], c["n"])
        got = aqe.reference_order(ids, [tuple(o) for o in c["ops"]])
        assert hashlib.sha256(got.astype("<i8").tobytes()).hexdigest() == c["sha256"], c["name"]
        assert bool((np.diff(ids[got.astype(np.int64)]) >= 0).all()) == c["sorted_by_id"], c["name"]


def test_order_without_duplicates_is_ascending_id_whatever_the_history():
    rng = np.random.default_rng(5)
    ids = rng.permutation(20000)
    for ops in (None, [(12000, 0), (8000, 0)], [(1, 0)] * 300 + [(19700, 0)]):
        got = aqe.reference_order(ids, ops)
        assert (ids[got.astype(np.int64)] == np.arange(20000)).all()


def test_restored_table_keeps_its_order_and_later_batches_go_where_the_tree_puts_them():
    """A table pulled back from the device before an append arrives as rows already in table order (kind 1)."""
    g = json.load(open(GOLDEN))
    c = [x for x in g["cases"] if x["name"] == "three_batches_append_only"][0]
    ids = np.asarray(c["ids"], dtype=np.int64)
    first = c["ops"][0][0]
    head = aqe.reference_order(ids[:first], [(first, 0)]).astype(np.int64)          # the table after the first batch ...
    ids2 = np.concatenate([ids[:first][head], ids[first:]])                          # ... restored in that order, then the other batches
    got = aqe.reference_order(ids2, [(first, 1)] + [tuple(o) for o in c["ops"][1:]]).astype(np.int64)
    arrival = np.concatenate([head, np.arange(first, len(ids))])                      # arrival numbers of ids2's rows in the original history
    assert [int(x) for x in arrival[got]] == c["order"]


def test_bad_histories_are_refused():
    with pytest.raises(RuntimeError):
        aqe.reference_order([1, 2, 3], [(2, 0)])
    with pytest.raises(RuntimeError):
        aqe.reference_order([1, 2, 3], [(3, 7)])
    assert len(aqe.reference_order([], None)) == 0


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libaqe_ref.so not built (no /root/reference here)")
def test_duplicate_id_order_matches_live_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_order_golden import reference_order as ref_order
    rng = np.random.default_rng(99)
    for trial in range(12):
        n = int(rng.integers(300, 6000))
        ids = rng.integers(0, int(rng.integers(2, 2000)), n)
        cuts = np.sort(rng.choice(np.arange(1, n), size=int(rng.integers(0, 4)), replace=False)) if n > 8 else []
        ops, at = [], 0
        for cut in list(cuts) + [n]:
            ops.append((int(cut - at), 0)); at = int(cut)
        if trial % 3 == 0:                      # a tail of single insert_record calls
            k = min(40, ops[-1][0] - 1)
            ops[-1] = (ops[-1][0] - k, 0)
            ops += [(1, 0)] * k
        want = ref_order(np.asarray(ids, dtype=np.int64), ops)
        got = aqe.reference_order(ids, ops).astype(np.int64)
        assert np.array_equal(got, want), (trial, n, ops[:5])
