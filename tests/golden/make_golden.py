#!/usr/bin/env python
"""Regenerates tests/golden/synth_*.{json,cmr.txt,coords.txt,lp.txt,lp_maximal.txt} by running the REFERENCE ITSELF
(oracle/_ref, compiled from /root/reference by oracle/Makefile) on seeded synthetic inputs made by
pacbio_b200/tools/gen_synth.  Run in the build container only (needs oracle/_ref):

    python tests/golden/make_golden.py

The fixtures pin the oracle port (tests/test_oracle.py) and, on the GPU, the CUDA path.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import REF_CMR, REF_JFA, REF_LP, Ref, gen_synth  # noqa: E402

CONFIGS = {
    "synth_g1": dict(gen=dict(genome=120000, coverage=3, read_len=4000, error=0.15, seed=11, repeat_frac=0.1),
                     mer=15, psa_min=13, unitig_k=41),
    "synth_g2": dict(gen=dict(genome=100000, coverage=3, read_len=5000, error=0.12, seed=12, repeat_frac=0.0),
                     mer=17, psa_min=13, unitig_k=41),
    "synth_g3": dict(gen=dict(genome=60000, coverage=4, read_len=3000, error=0.10, seed=13, repeat_frac=0.2,
                              mean_unitig=200),
                     mer=19, psa_min=10, unitig_k=31),
}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def queries(k, n, seed, text_codes):
    """Half sampled from the text (both strands), half uniform random."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 4 ** k, size=n, dtype=np.uint64)
    pos = rng.integers(0, len(text_codes) - k, size=n // 2)
    for i, p in enumerate(pos):
        v = 0
        codes = text_codes[p:p + k]
        if i & 1:
            codes = 3 - codes[::-1]
        for c in codes:
            v = (v << 2) | int(c)
        q[i] = v
    return q


def text_codes_of(sr_fasta):
    seq = []
    with open(sr_fasta) as f:
        for line in f:
            if not line.startswith(">"):
                seq.append(line.strip())
    b = np.frombuffer("".join(seq).encode(), dtype=np.uint8)
    return ((b >> 1) ^ (b >> 2)) & 3


LP_VARIANTS = {"lp": [], "lp_maximal": ["-T", "maximal", "--trim", "match", "-b", "-O", "1.5", "-d", "0.01"]}


def longest_path_goldens():
    """synth_*.lp*.txt: the reference's longest_path_overlap_graph2 on the committed coords files."""
    for name, cfg in CONFIGS.items():
        with tempfile.TemporaryDirectory() as tmp:
            info = gen_synth(os.path.join(tmp, name), **cfg["gen"])
            for tag, extra in LP_VARIANTS.items():
                subprocess.check_call([REF_LP, "-k", str(cfg["unitig_k"]), "-l", info["unitigs_len"], "-t", "1"] + extra +
                                      ["-o", os.path.join(HERE, "%s.%s.txt" % (name, tag)), os.path.join(HERE, name + ".coords.txt")])
            print(name, "longest path goldens written")


# -F (fine pass): which fine mers are pinned for which fixture, and whether the jf_aligner coords are kept too
FINE = {"synth_g1": [(11, True), (14, False)], "synth_g2": [(13, False)], "synth_g3": [(12, True)]}


def fine_goldens():
    """synth_*.fine<F>.{cmr,coords}.txt: the reference's create_mega_reads / jf_aligner with -F."""
    for name, cfg in CONFIGS.items():
        with tempfile.TemporaryDirectory() as tmp:
            info = gen_synth(os.path.join(tmp, name), **cfg["gen"])
            common = ["-s", "1M", "-m", str(cfg["mer"]), "--psa-min", str(cfg["psa_min"]), "-k", str(cfg["unitig_k"]),
                      "-l", info["unitigs_len"], "-r", info["sr"], "-p", info["reads"]]
            for fine, with_coords in FINE[name]:
                cmr = os.path.join(HERE, "%s.fine%d.cmr.txt" % (name, fine))
                subprocess.check_call([REF_CMR] + common + ["-F", str(fine), "-t", "1", "-o", cmr], stderr=subprocess.DEVNULL)
                if with_coords:
                    subprocess.check_call([REF_JFA] + common + ["-F", str(fine), "-t", "1", "-H", "--coords",
                                                                os.path.join(HERE, "%s.fine%d.coords.txt" % (name, fine))],
                                          stderr=subprocess.DEVNULL)
            print(name, "fine pass goldens written")


def main():
    if "--only-longest-path" in sys.argv:
        return longest_path_goldens()
    if "--only-fine" in sys.argv:
        return fine_goldens()
    ref = Ref()
    for name, cfg in CONFIGS.items():
        with tempfile.TemporaryDirectory() as tmp:
            info = gen_synth(os.path.join(tmp, name), **cfg["gen"])
            k, m, uk = cfg["mer"], cfg["psa_min"], cfg["unitig_k"]
            meta = dict(config=cfg, inputs={key: sha(open(info[key], "rb").read())
                                            for key in ("sr", "reads", "unitigs", "unitigs_len")})
            h = ref.index_create(info["sr"], m, k)
            sa = ref.sa(h)
            counts = ref.counts(h, m)
            meta["n"] = int(ref.n(h))
            meta["sa_sha256"] = sha(sa.astype("<u8").tobytes())
            meta["counts_sha256"] = sha(counts.astype("<u8").tobytes())
            q = queries(k, 4000, 99, text_codes_of(info["sr"]))
            idx, nb = ref.search(h, q)
            meta["search"] = dict(n=4000, seed=99, index_sha256=sha(idx.astype("<u8").tobytes()),
                                  nb_sha256=sha(nb.astype("<u8").tobytes()), nb_sum=int(nb.sum()),
                                  first=[[int(a), int(b), int(c)] for a, b, c in zip(q[:8], idx[:8], nb[:8])])
            ref.index_destroy(h)
            cmr = os.path.join(HERE, name + ".cmr.txt")
            subprocess.check_call([REF_CMR, "-s", "1M", "-m", str(k), "--psa-min", str(m), "-k", str(uk), "-l",
                                   info["unitigs_len"], "-t", "4", "-B", "17", "--max-count", "5000", "-d", "0.029",
                                   "-r", info["sr"], "-p", info["reads"], "-o", cmr], stderr=subprocess.DEVNULL)
            # canonical order: records sorted by header (thread order is arbitrary in the reference)
            recs, cur = [], None
            for line in open(cmr):
                if line.startswith(">"):
                    cur = [line]
                    recs.append(cur)
                else:
                    cur.append(line)
            recs.sort(key=lambda r: r[0])
            open(cmr, "w").write("".join("".join(r) for r in recs))
            cmr_u = os.path.join(tmp, "cmr_u.txt")
            subprocess.check_call([REF_CMR, "-s", "1M", "-m", str(k), "--psa-min", str(m), "-k", str(uk), "-u",
                                   info["unitigs"], "-t", "1", "-B", "17", "--max-count", "5000", "-d", "0.029",
                                   "-r", info["sr"], "-p", info["reads"], "-o", cmr_u], stderr=subprocess.DEVNULL)
            meta["cmr_with_sequences_sha256_t1"] = sha(open(cmr_u, "rb").read())
            coords = os.path.join(HERE, name + ".coords.txt")
            subprocess.check_call([REF_JFA, "-s", "1M", "-m", str(k), "--psa-min", str(m), "-k", str(uk), "-l",
                                   info["unitigs_len"], "-t", "1", "-H", "-r", info["sr"], "-p", info["reads"],
                                   "--coords", coords], stderr=subprocess.DEVNULL)
            json.dump(meta, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
            print(name, "n", meta["n"], "records", len(recs))


if __name__ == "__main__":
    main()
    if "--only-longest-path" not in sys.argv and "--only-fine" not in sys.argv:
        longest_path_goldens()
        fine_goldens()
