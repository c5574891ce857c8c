"""CPU tests pinning the oracle port (oracle/port) to (1) the reference's own golden vectors and
(2) outputs of the reference itself (tests/golden/synth_*, made by tests/golden/make_golden.py from
oracle/_ref) and, when oracle/_ref is present, live stage-by-stage comparisons."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle_lib import gen_synth, read_fasta, records

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
UNIT = json.load(open(os.path.join(GOLD, "unit_vectors.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


# ---- reference unit-test vectors ----------------------------------------------------------------
@pytest.mark.parametrize("case", UNIT["kmers_info"], ids=lambda c: c["name"])
def test_kmers_info_known_answers(port, case):
    rows = port.kmers_info_trace(case["name"], case["unitig_lengths"], case["unitigs_k"], case["k"],
                                 case["positions"])
    assert [r[0] for r in rows] == case["mers"]
    assert [r[1] for r in rows] == case["bases"]


@pytest.mark.parametrize("case", UNIT["lis"])
def test_lis_known_answers(port, case):
    res = port.lis(case["pairs"], case["a"], case["b"], case["cap"])
    if "expect" in case:
        assert res.tolist() == case["expect"]
    else:
        assert len(res) == case["expect_len"]
    sr = [case["pairs"][i][1] for i in res]
    assert all(x < y for x, y in zip(sr, sr[1:]))


def test_sr_overlap_known_answers(port):
    sr1, sr2, sr2r, sr3, sr4 = "1F_2R_3F_4R", "4R_5F_6R", "4F_5R_6F", "1F_2R_7F_1F_2R", "2R"
    assert port.sr_overlap("", "") == 0 and port.sr_overlap("", sr1) == 0 and port.sr_overlap(sr1, "") == 0
    assert port.sr_overlap(sr1, sr2) == 1 and port.sr_overlap(sr2, sr1) == 0
    assert port.sr_overlap(sr1, sr2r) == 0 and port.sr_overlap(sr2r, sr1) == 0
    assert port.sr_overlap(sr3, sr1) == 2 and port.sr_overlap(sr3, sr4) == 0 and port.sr_overlap(sr3, sr2) == 0


# ---- reference CLI goldens (tests/aligner_output; k=17, --stretch-cap 200) ------------------------
def _golden_coords(path):
    """The reference's golden files use the old non-compact layout `... Err Rname Qname [info]`."""
    rows = []
    for line in open(path).read().splitlines()[1:]:
        f = line.split()
        rows.append(tuple(f[:14]) + tuple(f[15:]))
    return sorted(rows)


@pytest.mark.parametrize("forward", [False, True])
def test_aligner_output_goldens(port, tmp_path, forward):
    d = os.path.join(GOLD, "aligner_output")
    out = str(tmp_path / "coords")
    port.run(1, os.path.join(d, "test_super_reads.fa"), os.path.join(d, "test_pacbio.fa"),
             os.path.join(d, "test_unitigs_lengths") if forward else None, out, 17, 65 if forward else 0,
             unitigs_is_fasta=False, stretch_cap=200.0, forward=forward)
    got = sorted(tuple(l.split()) for l in open(out).read().splitlines() if not l.startswith(">"))
    assert got == _golden_coords(os.path.join(d, "coords_forward_expected" if forward else "coords_normal_expected"))


# ---- fixtures generated from the reference itself ---------------------------------------------------
@pytest.fixture(scope="module", params=["synth_g1", "synth_g2", "synth_g3"])
def golden_case(request, tmpdir_session):
    name = request.param
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    info = gen_synth(os.path.join(tmpdir_session, name), **meta["config"]["gen"])
    for key, h in meta["inputs"].items():
        assert sha(open(info[key], "rb").read()) == h, "generator output drifted for " + key
    return name, meta, info


def test_port_index_matches_reference_fixture(port, golden_case):
    from golden.make_golden import queries, text_codes_of
    name, meta, info = golden_case
    cfg = meta["config"]
    h = port.index_create(info["sr"], cfg["psa_min"], cfg["mer"])
    assert port.n(h) == meta["n"]
    assert sha(port.sa(h).astype("<u8").tobytes()) == meta["sa_sha256"]
    assert sha(port.counts(h, cfg["psa_min"]).astype("<u8").tobytes()) == meta["counts_sha256"]
    s = meta["search"]
    q = queries(cfg["mer"], s["n"], s["seed"], text_codes_of(info["sr"]))
    idx, nb = port.search(h, q)
    assert [[int(a), int(b), int(c)] for a, b, c in zip(q[:8], idx[:8], nb[:8])] == s["first"]
    assert int(nb.sum()) == s["nb_sum"]
    assert sha(idx.astype("<u8").tobytes()) == s["index_sha256"]
    assert sha(nb.astype("<u8").tobytes()) == s["nb_sha256"]
    port.index_destroy(h)


def test_port_text_matches_reference_fixture(port, golden_case, tmp_path):
    name, meta, info = golden_case
    cfg = meta["config"]
    out = str(tmp_path / "cmr.txt")
    port.run(0, info["sr"], info["reads"], info["unitigs_len"], out, cfg["mer"], cfg["unitig_k"],
             unitigs_is_fasta=False, psa_min=cfg["psa_min"], threads=2)
    assert records(out) == records(os.path.join(GOLD, name + ".cmr.txt"))
    out = str(tmp_path / "coords.txt")
    port.run(1, info["sr"], info["reads"], info["unitigs_len"], out, cfg["mer"], cfg["unitig_k"],
             unitigs_is_fasta=False, psa_min=cfg["psa_min"], threads=2)
    assert records(out) == records(os.path.join(GOLD, name + ".coords.txt"))
    # with sequences (-u): single thread keeps the reference's record order, so whole-file hash
    out = str(tmp_path / "cmr_u.txt")
    port.run(0, info["sr"], info["reads"], info["unitigs"], out, cfg["mer"], cfg["unitig_k"],
             unitigs_is_fasta=True, psa_min=cfg["psa_min"], threads=1)
    assert sha(open(out, "rb").read()) == meta["cmr_with_sequences_sha256_t1"]


FINE = {"synth_g1": [(11, True), (14, False)], "synth_g2": [(13, False)], "synth_g3": [(12, True)]}   # as in make_golden.py


def test_port_fine_pass_matches_reference_fixture(port, golden_case, tmp_path):
    """-F: the port's restatement of fine_aligner.cc against the reference's own output (fine mer below,
    at and above --psa-min)."""
    name, meta, info = golden_case
    cfg = meta["config"]
    for fine, with_coords in FINE[name]:
        out = str(tmp_path / "cmr.txt")
        port.run(0, info["sr"], info["reads"], info["unitigs_len"], out, cfg["mer"], cfg["unitig_k"],
                 unitigs_is_fasta=False, psa_min=cfg["psa_min"], threads=2, fine_mer=fine)
        assert records(out) == records(os.path.join(GOLD, "%s.fine%d.cmr.txt" % (name, fine)))
        if with_coords:
            out = str(tmp_path / "coords.txt")
            port.run(1, info["sr"], info["reads"], info["unitigs_len"], out, cfg["mer"], cfg["unitig_k"],
                     unitigs_is_fasta=False, psa_min=cfg["psa_min"], threads=2, fine_mer=fine)
            assert records(out) == records(os.path.join(GOLD, "%s.fine%d.coords.txt" % (name, fine)))


# ---- live comparison against the compiled reference -------------------------------------------------
def _canon_coords(r):
    rows = []
    for j in range(len(r["cint"])):
        lo, hi = r["info_off"][j], r["info_off"][j + 1]
        rows.append((tuple(r["cint"][j]), tuple(r["cdbl"][j]), tuple(r["kinfo"][lo:hi]), tuple(r["binfo"][lo:hi])))
    return sorted(rows)


@pytest.mark.parametrize("k,max_match", [(15, False), (17, True)])
def test_port_stages_match_live_reference(port, ref, tmpdir_session, k, max_match):
    info = gen_synth(os.path.join(tmpdir_session, "live%d" % k), 150000, coverage=3, read_len=4000, seed=20 + k,
                     repeat_frac=0.15)
    hp, hr = port.index_create(info["sr"], 13, k), ref.index_create(info["sr"], 13, k)
    assert np.array_equal(port.sa(hp), ref.sa(hr))
    assert np.array_equal(port.counts(hp, 13), ref.counts(hr, 13))
    ul = np.loadtxt(info["unitigs_len"], dtype=np.int64)[:, 1]
    port.set_unitigs_lengths(hp, ul)
    ref.set_unitigs_lengths(hr, ul)
    ap = port.aligner_create(hp, unitigs_k=41, max_match=max_match)
    ar = ref.aligner_create(hr, unitigs_k=41, max_match=max_match)
    _, seqs = read_fasta(info["reads"])
    for s in seqs[:40]:
        a, b = port.align_read(ap, s), ref.align_read(ar, s)
        for key in ("groups", "offsets", "lis"):
            assert np.array_equal(a[key], b[key]), key
        assert _canon_coords(a) == _canon_coords(b)     # doubles compared bit for bit


def test_lis_random_vs_live_reference(port, ref):
    rng = np.random.default_rng(5)
    for t in range(300):
        n = int(rng.integers(1, 150))
        pb = np.sort(rng.integers(1, 400, size=n))
        sr = rng.integers(1, 400, size=n) if t % 3 else pb + rng.integers(-30, 30, size=n)
        pairs = np.stack([pb, sr], axis=1)
        assert port.lis(pairs).tolist() == ref.lis(pairs).tolist()
        for window in (2, 3, 5):                        # --window-size (lis_align.hpp:17-45): the port's ring restatement
            assert port.lis(pairs, window=window).tolist() == ref.lis(pairs, window=window).tolist(), (t, window)
