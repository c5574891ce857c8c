#!/bin/bash
# BASELINE.json configs[3] shape on ONE GPU: synthetic human-size genome (3.1 Gbp, 20 % repeats),
# super-reads of > 2^32 bases (an index of several parts), a thin slice of 15 kbp reads at 15 % error.
# Checks that the records do not depend on where the super-reads are cut (2 parts vs 3 parts) and,
# when oracle/_ref is there and REF=1, compares with the reference binary.  Writes a summary on stdout.
set -u
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
D=${MR_HUMAN_DIR:-/tmp/mr_human}; mkdir -p $D
GENOME=${GENOME:-3100000000}; COV=${COV:-0.02}; SRCOV=${SRCOV:-1.4}
SECONDS=0
[ -x $ROOT/pacbio_b200/tools/gen_synth ] || g++ -O2 -std=c++17 -pthread $ROOT/pacbio_b200/tools/gen_synth.cc -o $ROOT/pacbio_b200/tools/gen_synth
$ROOT/pacbio_b200/tools/gen_synth --genome $GENOME --coverage $COV --read-len 15000 --error 0.15 --seed 45 --sr-cov $SRCOV \
   --repeat-frac 0.2 --unitig-k 41 --threads $(nproc) --prefix $D/h > $D/gen.json || exit 1
echo "generated: $(cat $D/gen.json) in $SECONDS s"
ARGS="-s 1M -m 15 --psa-min 13 --stretch-cap 10000 -k 41 -l $D/h.unitigs_len.txt -B 17 -d 0.029 --max-count 5000 -t $(nproc) -r $D/h.superreads.fa -p $D/h.reads.fa"
MR_SHOW_TIMING=1 $ROOT/pacbio_b200/bin/create_mega_reads $ARGS -o $D/a.txt 2> $D/a.err; echo "run A (default cut) rc=$?"; cat $D/a.err
MR_SHOW_TIMING=1 MR_INDEX_PART_BASES=${PART_B:-1600000000} $ROOT/pacbio_b200/bin/create_mega_reads $ARGS -o $D/b.txt 2> $D/b.err; echo "run B (smaller parts) rc=$?"; cat $D/b.err
if cmp -s $D/a.txt $D/b.txt; then echo "A == B: $(wc -l < $D/a.txt) lines, $(grep -c '^>' $D/a.txt) reads with mega-reads, sha256 $(sha256sum < $D/a.txt | cut -c1-16)"; else echo "A != B"; fi
if [ "${REF:-0}" = 1 ] && [ -x $ROOT/oracle/_ref/create_mega_reads ]; then
  SECONDS=0
  timeout ${REF_TIMEOUT:-500} $ROOT/oracle/_ref/create_mega_reads $ARGS -o $D/ref.txt 2> $D/ref.err; rc=$?
  echo "reference rc=$rc in $SECONDS s"; tail -n 5 $D/ref.err
  if [ $rc = 0 ]; then
    # our coords rows, to tell which reads hold an exact (rs, re, ql) tie: there the reference's own
    # record is not defined (unordered_map pointer order + unstable sort, create_mega_reads.cc:69-77)
    $ROOT/pacbio_b200/bin/jf_aligner -s 1M -m 15 --psa-min 13 --stretch-cap 10000 -k 41 -l $D/h.unitigs_len.txt -B 17 --max-count 5000 \
        -H --coords $D/a.coords -r $D/h.superreads.fa -p $D/h.reads.fa 2> $D/jfa.err || cat $D/jfa.err
    python - <<PY
import sys, json
sys.path.insert(0, "$ROOT/tests")
from oracle_lib import records
a, b = records("$D/a.txt"), records("$D/ref.txt")
diff = sorted(k for k in set(a) | set(b) if a.get(k) != b.get(k))
ties, cur, seen = set(), None, None
for line in open("$D/a.coords"):
    if line.startswith(">"):
        cur, seen = line.split()[1], set()
    else:
        f = line.split()
        key = (f[0], f[1], f[10])
        if key in seen:
            ties.add(cur)
        seen.add(key)
hard = [k for k in diff if k[1:] not in ties]
print("reference comparison: %d records, %d differ, %d of them on reads without an exact (rs, re, ql) coords tie; reads with such ties: %d"
      % (len(b), len(diff), len(hard), len(ties)))
out = "${MR_HUMAN_OUT:-$D}/human_scale_diff.json"
json.dump({k: {"ours": a.get(k), "reference": b.get(k), "tie": k[1:] in ties} for k in diff}, open(out, "w"), indent=1)
PY
  fi
fi
