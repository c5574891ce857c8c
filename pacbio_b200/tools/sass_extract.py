#!/usr/bin/env python
"""Cuts the SASS of selected kernels out of `cuobjdump -sass libmegareads_b200.so` and summarises the instruction mix:
    python pacbio_b200/tools/sass_extract.py <name-substring> [<name-substring> ...] > listing.txt
Every listing starts with the mnemonic histogram of the kernel (the memory-movement ones first: UBLKCP = bulk
asynchronous copy issued to the TMA engine, SYNCS = mbarrier operations, LDGSTS = cp.async, LDG/STG/LDS/STS/ATOMS,
MATCH/VOTE/SHFL/REDUX warp collectives, DFMA/DADD/DMUL FP64), then the full SASS."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FIRST = ("UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "MATCH", "VOTE", "SHFL", "REDUX",
         "DFMA", "DADD", "DMUL", "DSETP", "BAR", "FENCE")


def main():
    text = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "pacbio_b200", "libmegareads_b200.so")],
                          capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"(?m)^\s*Function : ", text)[1:]
    for want in sys.argv[1:]:
        for f in funcs:
            name = f.split("\n", 1)[0].strip()
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            if want not in demangled:
                continue
            ops = collections.Counter()
            for line in f.split("\n"):
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
                if m:
                    ops[m.group(1).split(".")[0]] += 1
            total = sum(ops.values())
            print("=" * 100)
            print("kernel:", demangled)
            print("instructions:", total)
            print("memory movement / collectives / FP64:", ", ".join("%s %d" % (k, ops[k]) for k in FIRST if ops.get(k)))
            print("other:", ", ".join("%s %d" % (k, v) for k, v in ops.most_common() if k not in FIRST))
            print("=" * 100)
            print("        Function : " + f.rstrip())
            print()


if __name__ == "__main__":
    main()
