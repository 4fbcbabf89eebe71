"""Per-kernel resources and SASS evidence of the built library, from the build's own logs.

    python scripts/sass_resources.py > profiles/r02_sass_resources.md

Reads ``ser_b200/_build/ptxas.log`` (``-Xptxas -v`` of every source, written by
``ser_b200/build.py``) and ``cuobjdump -sass ser_b200/libser_b200.so``; needs no GPU.  For
every ``__global__`` function: registers, static shared memory, spill bytes, and how often the
sm_100a-specific instructions occur in its SASS -- ``UBLKCP`` (cp.async.bulk, the TMA bulk copy of
the staging paths), ``UTCHMMA`` / ``LDTM`` / ``UTCBAR`` (tcgen05.mma / tcgen05.ld / tcgen05.commit),
``SYNCS`` (mbarrier), packed ``FFMA2 / FADD2 / FMUL2``, ``LDGSTS`` (cp.async), ``FMNMX`` (medians).
"""

from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
LIB = REPO / "ser_b200" / "libser_b200.so"
PTXAS_LOG = REPO / "ser_b200" / "_build" / "ptxas.log"

MNEMONICS = ("UBLKCP", "UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX", "DFMA")


def demangle(names: list[str]) -> dict[str, str]:
    out = subprocess.run(["c++filt", *names], capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name: str) -> str:
    """`void ser::cqtc_kernel<16, true>(ser::CqtcArgs)` -> `cqtc_kernel<16, true>`"""
    name = re.sub(r"^void\s+", "", name)
    depth = 0
    for i, ch in enumerate(name):          # cut the parameter list: first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            name = name[:i]
            break
    return re.sub(r"^((\w+|\(anonymous namespace\))::)+", "", name)


def ptxas_table() -> "OrderedDict[str, dict]":
    table: OrderedDict[str, dict] = OrderedDict()
    source = "?"
    current = None
    for line in PTXAS_LOG.read_text().splitlines():
        if line.startswith("--- "):
            source = line[4:].strip()
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            current = table.setdefault(m.group(1), {"source": source})
            continue
        if current is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            current.update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            current["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            current["smem"] = int(s.group(1)) if s else 0
            b = re.search(r"used (\d+) barriers", line)
            current["barriers"] = int(b.group(1)) if b else 0
            current = None
    return table


def sass_counts() -> dict[str, Counter]:
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    counts: dict[str, Counter] = {}
    fn = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = Counter()
            continue
        if fn is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[fn]["_all"] += 1
            for mn in MNEMONICS:
                if op == mn:
                    counts[fn][mn] += 1
    return counts


def main() -> int:
    if not LIB.exists() or not PTXAS_LOG.exists():
        print("build first: python -m ser_b200.build --force", file=sys.stderr)
        return 1
    table = ptxas_table()
    counts = sass_counts()
    names = demangle(list(table))
    print("# Resources and sm_100a instructions per kernel (`scripts/sass_resources.py`)\n")
    print("`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`, CUDA 12.9; registers / static shared")
    print("memory / spills from `-Xptxas -v`, instruction counts from `cuobjdump -sass` of the linked library.")
    print("Dynamic shared memory is set at launch (see `csrc/api.cu`) and is not in this table.\n")
    head = ["kernel", "source", "regs", "static smem B", "spill st/ld B", "SASS instr"] + list(MNEMONICS)
    print("| " + " | ".join(head) + " |")
    print("|" + "---|" * len(head))
    totals = Counter()
    for mangled, row in table.items():
        c = counts.get(mangled, Counter())
        totals.update(c)
        cells = [f"`{short(names.get(mangled, mangled))}`", row["source"], str(row.get("regs", "?")),
                 str(row.get("smem", 0)), f"{row.get('spill_st', 0)}/{row.get('spill_ld', 0)}", str(c["_all"])]
        cells += [str(c[mn]) if c[mn] else "" for mn in MNEMONICS]
        print("| " + " | ".join(cells) + " |")
    print("| **total** | | | | | " + str(totals["_all"]) + " | " + " | ".join(str(totals[mn]) for mn in MNEMONICS) + " |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
