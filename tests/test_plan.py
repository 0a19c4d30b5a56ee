"""Host-side symbolic plan: bit-exact with the reference (no GPU needed).

Rows A0-A5 of SURVEY.md section 8a.  Golden sources:
  * the FNV hashes SURVEY.md section 8c quotes for P / Parent / Lp / Li of the DogBot patterns,
  * tests/golden/plans.npz and amd_cases.npz, produced by tests/golden/make_golden.py from the
    reference compiled from /root/reference,
  * the live reference (oracle/_ref) when present.
"""
import os

import numpy as np
import pytest

from apf_quadruped_b200 import dogbot

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# SURVEY.md section 8c (word-wise FNV-1a-64 of the long arrays, reference-computed)
SURVEY_HASHES = {
    "stance": ("ca5ca33732f4a1af", "e3366024eef860c2", "e3823bfdd73f9bbd", "0d5ef56ca82ac26f"),
    "crawl": ("745292d5a5cb7c7e", "1082d3571f63ee7c", "8834f7f2cada3ea9", "7573429dca22ac74"),
    "trot": ("57d51943f7b65b03", "55728fdd05006f80", "7fd2270d169b8446", "9be954a737ff42cf"),
}
# n, p, m, nnzP, nnzA, nnzG, N, nnzK, nnzU, Lnz, max col, factor flops
SURVEY_TABLE = {
    "stance": (30, 18, 68, 162, 324, 636, 116, 2150, 1124, 1686, 47, 49842),
    "crawl": (30, 15, 69, 102, 252, 669, 114, 2013, 1056, 1485, 41, 37853),
    "trot": (30, 12, 70, 60, 180, 702, 112, 1894, 997, 1179, 23, 22053),
    "mpc": (240, 120, 240, 240, 660, 400, 600, 2600, 1540, 2405, 22, 24385),
}


def fnv(a):
    h = 1469598103934665603
    for v in np.asarray(a, dtype=np.int64).tolist():
        h = ((h ^ (v & 0xFFFFFFFFFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def test_library_exports_every_declared_symbol(qps):
    lb = qps.lib()
    for name in qps.EXPORTS:
        assert hasattr(lb, name), name
    # every function include/*.h declares is in EXPORTS
    import re
    root = os.path.dirname(os.path.dirname(__file__))
    declared = set()
    for hdr in ("include/qpSWIFT/qpSWIFT.h", "include/qpswift_b200.h"):
        text = open(os.path.join(root, hdr)).read()
        declared |= set(re.findall(r"\b(QPB?_[A-Za-z_]+)\s*\(", text))
    declared = {d for d in declared if not d.startswith("QPB_E_")} - {"QP_OPTIMAL", "QP_KKTFAIL", "QP_MAXIT", "QP_FATAL"}
    assert declared <= set(qps.EXPORTS), declared - set(qps.EXPORTS)


@pytest.mark.parametrize("kind", list(SURVEY_TABLE))
def test_pattern_sizes_match_survey(qps, kind):
    pat = dogbot.pattern(kind)
    plan = qps.Plan(pat)
    n, p, m, nP, nA, nG, N, nK, nU, lnz, maxcol, flops = SURVEY_TABLE[kind]
    assert (pat.n, pat.p, pat.m, pat.nnzP, pat.nnzA, pat.nnzG) == (n, p, m, nP, nA, nG)
    assert plan.info("N") == N and plan.info("nnzK") == nK and plan.info("nnzU") == nU
    assert plan.info("lnz") == lnz and plan.info("factor_flops") == flops
    assert int(plan.array("Lnz").max()) == maxcol
    assert plan.info("amd_status") == 0 and plan.info("valid") == 1


@pytest.mark.parametrize("kind", list(SURVEY_HASHES))
def test_plan_hashes_match_survey(qps, kind):
    plan = qps.Plan(dogbot.pattern(kind))
    got = tuple(fnv(plan.array(k)) for k in ("P", "Parent", "Lp", "Li"))
    assert got == SURVEY_HASHES[kind]


def test_stance_permutation_prefix_suffix(qps):
    P = qps.Plan(dogbot.pattern("stance")).array("P")
    assert P[:16].tolist() == [50, 48, 52, 51, 49, 55, 53, 57, 56, 54, 60, 58, 62, 61, 59, 65]
    assert P[-4:].tolist() == [89, 90, 91, 18]


@pytest.mark.parametrize("kind", list(SURVEY_TABLE))
def test_plan_matches_golden(qps, kind):
    gold = np.load(os.path.join(GOLD, "plans.npz"))
    plan = qps.Plan(dogbot.pattern(kind))
    for k in ("Kjc", "Kir", "P", "Pinv", "Parent", "Lnz", "Lp", "Li"):
        assert np.array_equal(plan.array(k), gold[f"{kind}_{k}"]), k


def test_amd_matches_golden_cases(qps):
    g = np.load(os.path.join(GOLD, "amd_cases.npz"))
    jumbled = 0
    for t, n in enumerate(g["n"].tolist()):
        Ap = g["Ap"][g["Ap_off"][t]:g["Ap_off"][t + 1]]
        Ai = g["Ai"][g["Ai_off"][t]:g["Ai_off"][t + 1]]
        P = g["P"][g["P_off"][t]:g["P_off"][t + 1]]
        st, mine = qps.amd_order(n, Ap, Ai)
        assert st == int(g["status"][t]), t
        assert np.array_equal(mine, P), t
        jumbled += st == 1
    assert jumbled > 10  # the unsorted-input (AMD_preprocess) path is exercised


def test_amd_matches_live_reference(qps, oracle):
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    import scipy.sparse as sp
    rng = np.random.default_rng(99)
    for t in range(300):
        n = int(rng.integers(1, 160))
        A = sp.random(n, n, density=float(rng.choice([0.01, 0.05, 0.2, 0.6])), random_state=int(rng.integers(1 << 31)), format="csc")
        if t % 3 == 0:
            A = (A + A.T).tocsc()
        if t % 11 == 0:
            A = A.tolil()
            A[0, :] = 1
            A = A.tocsc()
        A.sort_indices()
        Ap, Ai = A.indptr.astype(np.int64), A.indices.astype(np.int64)
        assert qps.amd_order(n, Ap, Ai)[1].tolist() == oracle.amd_order(n, Ap, Ai)[1].tolist(), t


def test_amd_rejects_invalid(qps):
    st, _ = qps.amd_order(3, np.array([0, 1, 2, 3]), np.array([0, 5, 2]))
    assert st == -2
    st, _ = qps.amd_order(3, np.array([1, 1, 2, 3]), np.array([0, 1, 2]))
    assert st == -2


def test_user_permutation_is_kept(qps):
    pat = dogbot.pattern("trot")
    perm = np.random.default_rng(1).permutation(pat.N)
    plan = qps.Plan(pat, perm)
    assert np.array_equal(plan.array("P"), perm)
    assert plan.info("amd_status") == -3
    assert np.array_equal(plan.array("Pinv")[perm], np.arange(pat.N))


def test_empty_inequality_row_is_refused(qps):
    pat = dogbot.pattern("trot")
    keep = pat.Gir != 3
    Gir = pat.Gir[keep]
    Gjc = np.concatenate([[0], np.cumsum([np.sum(keep[pat.Gjc[c]:pat.Gjc[c + 1]]) for c in range(pat.n)])]).astype(np.int64)
    bad = dogbot.Pattern("bad", pat.n, pat.m, pat.p, pat.Pjc, pat.Pir, pat.Ajc, pat.Air, Gjc, Gir)
    assert qps.Plan(bad).info("valid") == 0


def test_generator_is_shard_independent():
    whole = dogbot.generate("crawl", 2500, start=100)
    a = dogbot.generate("crawl", 1000, start=100)
    b = dogbot.generate("crawl", 1500, start=1100)
    for k in ("Ppr", "Apr", "Gpr", "c", "h", "b"):
        assert np.array_equal(getattr(whole, k), np.concatenate([getattr(a, k), getattr(b, k)]))
