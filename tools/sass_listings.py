#!/usr/bin/env python
"""Regenerate profiles/sass_r2: full `cuobjdump -sass` listings (gzip) of the tensor-core / collective objects of the
CURRENT build and a mnemonic histogram per object (tcgen05 = UTC*MMA, TMEM loads = LDTM, TMA = UTMA*, multimem = *MC*).

    python tools/sass_listings.py            # after __graft_entry__.build()
"""
import collections
import gzip
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "tnn_b200" / "_build"
OUT = ROOT / "profiles" / "sass_r2"
OBJECTS = ["umma_gemm", "mxfp8", "gemm_xt", "attention", "attention_bwd", "decode_mega", "skinny_gemm", "norm", "optim"]
KEEP = re.compile(r"^(UTC|UTMA|LDTM|STTM|SYNCS|UBLKCP|LDGMC|REDG|RED\b|ATOMS|ATOMG|MEMBAR|REDUX|HMMA|DFMA|UTCBAR|LDGSTS|UCGABAR|ERRBAR|ACQBULK)")


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    lines = ["# SASS mnemonic histograms of the round-2 final objects (full listings: profiles/sass_r2/*.sass.gz, "
             "`cuobjdump -sass tnn_b200/_build/<name>.cu.o`; regenerate with `python tools/sass_listings.py`)\n"]
    for name in OBJECTS:
        obj = BUILD / f"{name}.cu.o"
        if not obj.exists():
            print(f"missing {obj}", file=sys.stderr)
            continue
        sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
        with gzip.open(OUT / f"{name}.sass.gz", "wt") as f:
            f.write(sass)
        hist = collections.Counter()
        total = 0
        for ln in sass.splitlines():
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
            if not m:
                continue
            total += 1
            op = m.group(1)
            if KEEP.match(op):
                hist[op] += 1
        lines.append(f"\n## {name}.cu.o  ({total} instructions)\n\n| mnemonic | count |\n|---|---|")
        for op, c in hist.most_common(24):
            lines.append(f"| `{op}` | {c} |")
    (OUT / "README.md").write_text("\n".join(lines) + "\n")
    print(f"wrote {OUT}")


if __name__ == "__main__":
    main()
