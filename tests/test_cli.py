"""The repaired command line (approximatequeryengine_b200/cli.py; the reference's is enhanced_aqe_cli.py): host logic on
the CPU, the three query syntaxes end to end on the GPU."""
import json
import math

import pytest

from approximatequeryengine_b200 import cli


def test_embedded_approx_and_aggregate_parsing():
    q, emb = cli.parse_embedded_approx("SELECT APPROX(SUM(amount)) FROM sales")         # enhanced_aqe_cli.py:83-95
    assert emb and q == "SELECT SUM(amount) FROM sales"
    q, emb = cli.parse_embedded_approx("select approx( AVG(amount) ) from sales where region = 1")
    assert emb and cli.aggregate_of(q) == ("AVG", "amount") and not cli.plain_amount_query(q)
    assert cli.parse_embedded_approx("SELECT SUM(amount) FROM sales") == ("SELECT SUM(amount) FROM sales", False)
    assert cli.plain_amount_query("SELECT COUNT(*) FROM sales") and cli.plain_amount_query("SELECT SUM(amount) FROM sales;")
    assert not cli.plain_amount_query("SELECT SUM(timestamp) FROM sales") and not cli.plain_amount_query("SELECT SUM(amount) FROM sales GROUP BY region")
    with pytest.raises(ValueError):
        cli.aggregate_of("SELECT MAX(amount) FROM sales")


def test_flags_the_reference_documents_are_accepted():
    # the reference's `--e` is ambiguous with `--explain` and `--s` never reaches a sampler; here both parse
    # (the call then fails on the missing file, before any device work)
    import contextlib
    import io
    q = "SELECT SUM(amount) FROM sales"
    for argv in ([q, "--s", "10"], [q, "-s", "10"], [q, "--sample", "10"], [q, "--e", "2"], [q, "-e", "2"], [q, "--error", "2", "--method", "clt"]):
        with contextlib.redirect_stderr(io.StringIO()) as err:
            rc = cli.main(argv + ["--db", "/nonexistent/file.aqe"])
        assert rc == 1 and "Could not open database" in err.getvalue(), (argv, err.getvalue())


def test_estimators_match_the_oracle(oracle):
    rows = oracle.synth(50000, seed=3)
    from oracle import make_params
    idx = oracle.indices(rows, "memory_stride", make_params("memory_stride", 2.0))
    s = oracle.stats(rows, idx)
    st = {"n": s.n, "mean": s.mean, "m2": s.m2, "sum": s.sum}
    for agg in ("SUM", "AVG", "COUNT"):
        v, m = cli.estimate_from_moments(agg, st, len(rows), 1.96)
        e, lo, hi = oracle.estimate(s, len(rows), agg.lower(), 1.96, legacy_ci=False)
        assert v == e and abs((hi - lo) / 2 - m) <= 1e-12 * max(m, 1.0)


@pytest.mark.gpu
def test_three_syntaxes_end_to_end(oracle, tmp_path, capsys):
    rows = oracle.synth(200000, seed=7)
    path = str(tmp_path / "sales.aqe")
    oracle.save_file(path, rows)
    exact = math.fsum(rows["amount"])

    def run(*argv):
        assert cli.main(list(argv) + ["--db", path, "--json"]) == 0
        return json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    r = run("SELECT SUM(amount) FROM sales")
    assert r["mode"] == "exact" and abs(r["value"] - exact) <= 1e-12 * exact
    r = run("SELECT APPROX(SUM(amount)) FROM sales", "--compare")
    assert r["mode"].startswith("memory_stride 10") and r["samples_used"] == 20000 and r["ci"][0] < exact < r["ci"][1] and r["actual_error_percent"] < 2
    r = run("SELECT SUM(amount) FROM sales", "--s", "1")
    from oracle import make_params
    idx = oracle.indices(rows, "memory_stride", make_params("memory_stride", 1.0))
    want = oracle.estimate(oracle.stats(rows, idx), len(rows), "sum")[0]
    assert abs(r["value"] - want) <= 1e-12 * want                       # the reference's E1 estimate on the reference's sample
    r = run("SELECT AVG(amount) FROM sales", "--e", "1")
    assert r["mode"].startswith("clt srs") and r["status"] == "STABLE" and abs(r["value"] - exact / len(rows)) / (exact / len(rows)) < 0.03
    r = run("SELECT SUM(amount) FROM sales", "--e", "1", "--method", "parallel")
    assert r["mode"].startswith("clt_validated_dual_pointer 20") and abs(r["value"] - exact) / exact < 0.03
    r = run("SELECT SUM(amount) FROM sales GROUP BY region", "--s", "10")
    assert set(r["value"]) == {str(k) for k in range(8)} and all(r["ci"][k][0] < r["value"][k] < r["ci"][k][1] for k in r["value"])
    r = run("SELECT COUNT(amount) FROM sales WHERE amount > 900")
    assert r["value"] == float((rows["amount"] > 900).sum())
