#!/usr/bin/env python
"""Pinned digests of the shishua byte stream (mitigation 3).

The reference clones shishua unpinned from GitHub at build time (Makefile.am:87-89) and the
source is not in this container, so `oracle/shishua.h` restates the published algorithm and the
raw stream is unpinned against UPSTREAM (it is bit-exact between oracle and GPU).  This script
records SHA-256 of the first 1 MiB of the stream for the seeds of the reference's first three
OpenMP threads ({1<<tid, 0, 0, 0}, src/HSimulation.tpp:28, src/inc/RNG.h:24):

    python tests/golden/make_shishua_digest.py            # (re)write shishua_sha256.json
    python tests/golden/make_shishua_digest.py --verify-upstream   # needs network + a C compiler

--verify-upstream clones github.com/espadrine/shishua, builds a 10-line program around its
shishua.h (prng_init / prng_gen, exactly the calls at src/RNG.cpp:24,29) and compares the
digests: one command closes the provenance gap on any machine with a network.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
OUT = os.path.join(HERE, "shishua_sha256.json")
SEEDS = [(1, 0, 0, 0), (2, 0, 0, 0), (4, 0, 0, 0)]
N_BYTES = 1 << 20

C_PROGRAM = r"""
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "shishua.h"
int main(int argc, char **argv) {
  uint64_t seed[4] = {strtoull(argv[1], 0, 10), 0, 0, 0};
  static uint8_t buf[1 << 20] __attribute__((aligned(128)));
  prng_state s;
  prng_init(&s, seed);
  prng_gen(&s, buf, sizeof buf);
  fwrite(buf, 1, sizeof buf, stdout);
  return 0;
}
"""


def oracle_digests():
    import oracle_api as oa
    return {",".join(map(str, s)): hashlib.sha256(oa.shishua_bytes(s, N_BYTES).tobytes()).hexdigest()
            for s in SEEDS}


def upstream_digests():
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["git", "clone", "--depth", "1", "https://github.com/espadrine/shishua", d + "/shishua"],
                       check=True)
        open(d + "/main.c", "w").write(C_PROGRAM)
        subprocess.run(["gcc", "-O2", "-march=native", "-I", d + "/shishua", "-o", d + "/gen", d + "/main.c"],
                       check=True)
        return {",".join(map(str, s)): hashlib.sha256(
            subprocess.run([d + "/gen", str(s[0])], check=True, capture_output=True).stdout).hexdigest()
            for s in SEEDS}


if __name__ == "__main__":
    mine = oracle_digests()
    if "--verify-upstream" in sys.argv:
        theirs = upstream_digests()
        for k in mine:
            print(k, "OK" if mine[k] == theirs[k] else f"MISMATCH oracle {mine[k]} upstream {theirs[k]}")
        sys.exit(0 if mine == theirs else 1)
    json.dump({"what": "SHA-256 of the first 1 MiB of the shishua stream (oracle/shishua.h), seed "
                       "{s0, s1, s2, s3} as key; NOT yet compared with upstream shishua "
                       "(run this script with --verify-upstream where a network exists)",
               "bytes": N_BYTES, "sha256": mine}, open(OUT, "w"), indent=1)
    print(open(OUT).read())
