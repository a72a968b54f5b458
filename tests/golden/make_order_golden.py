#!/usr/bin/env python3
"""Mints tests/golden/order_cases.json from the UNMODIFIED reference compiled in place (oracle/_ref/libaqe_ref.so, `make -C oracle ref`):
for tables with duplicate ids, the order in which the reference's B+ tree holds the rows after a history of insert_batch /
insert_record calls (read back through save_to_file = collect_all_records, custom_bplus_db.cpp:660-683).  Rows are told apart by
their timestamp field (= arrival number).  Run here (needs /root/reference); the JSON travels.

    python tests/golden/make_order_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import RECORD_DTYPE, Ref  # noqa: E402


def rows_of(ids):
    r = np.zeros(len(ids), dtype=RECORD_DTYPE)
    r["id"] = ids
    r["timestamp"] = np.arange(len(ids))
    r["amount"] = 1.0
    return r


def reference_order(ids, ops):
    """ops: [(rows, kind)], kind 0 = insert_batch of those rows (one row: insert_record), in arrival order."""
    R = Ref()
    rows = rows_of(ids)
    at = 0
    for cnt, kind in ops:
        assert kind == 0
        if cnt == 1:
            R.insert_record(rows[at:at + 1])
        else:
            R.insert_batch(rows[at:at + cnt])
        at += cnt
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "t.aqe")
        R.save(p)
        out = np.fromfile(p, dtype=RECORD_DTYPE, offset=24)
    assert len(out) == len(ids)
    return out["timestamp"].astype(np.int64)


def cases():
    rng = np.random.default_rng(20261019)
    out = []
    def add(name, ids, ops):
        out.append({"name": name, "ids": [int(x) for x in ids], "ops": [[int(a), int(b)] for a, b in ops]})
    add("tiny", [5, 3, 5, 5, 1, 3], [(6, 0)])
    add("all_equal_300", [7] * 300, [(300, 0)])
    add("all_equal_1000", [7] * 1000, [(1000, 0)])
    add("sorted_runs_2000", np.sort(rng.integers(0, 40, 2000)), [(2000, 0)])
    add("shuffled_runs_2000", rng.integers(0, 40, 2000), [(2000, 0)])
    add("few_dups_5000", np.where(rng.random(5000) < 0.02, 1234, rng.permutation(5000) + 10), [(5000, 0)])
    add("two_batches_interleaved", rng.integers(0, 300, 1500), [(900, 0), (600, 0)])
    add("three_batches_append_only", np.concatenate([rng.integers(0, 50, 700), rng.integers(50, 90, 800), rng.integers(90, 95, 400)]), [(700, 0), (800, 0), (400, 0)])
    add("singles_random_600", rng.integers(0, 60, 600), [(1, 0)] * 600)
    add("batch_then_singles", rng.integers(0, 100, 1200), [(1000, 0)] + [(1, 0)] * 200)
    return out


def main():
    cs = cases()
    for c in cs:
        c["order"] = [int(x) for x in reference_order(np.asarray(c["ids"], dtype=np.int64), c["ops"])]
    big = []   # too long to store element by element: ids = default_rng(seed).integers(0, hi, n); kept as a hash of the order
    import hashlib
    for name, seed, hi, n, ops in (("big_one_batch_80000", 71, 3000, 80000, [[80000, 0]]),
                                   ("big_two_batches_70000", 72, 20000, 70000, [[40000, 0], [30000, 0]]),
                                   ("big_singles_after_batch_45000", 73, 9000, 45000, [[40000, 0]] + [[1, 0]] * 5000)):
        ids = np.random.default_rng(seed).integers(0, hi, n)
        order = reference_order(np.asarray(ids, dtype=np.int64), ops)
        big.append({"name": name, "seed": seed, "hi": hi, "n": n, "ops": ops, "sorted_by_id": bool((np.diff(ids[order]) >= 0).all()),
                    "sha256": hashlib.sha256(order.astype("<i8").tobytes()).hexdigest()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "order_cases.json")
    with open(path, "w") as f:
        json.dump({"source": "oracle/_ref/libaqe_ref.so (unmodified custom_bplus_db.cpp), tests/golden/make_order_golden.py", "cases": cs, "big": big}, f)
    print("wrote", path, len(cs), "cases +", len(big), "hashed")


if __name__ == "__main__":
    main()
