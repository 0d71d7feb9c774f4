#!/usr/bin/env python3
"""SASS of the main loop of one kernel of libnttb200.so with its instruction mix.
Usage: python profiles/sass_loop.py <mangled kernel name> > profiles/<name>.txt   (no GPU needed)
The main loop is taken to be the span of the longest backward branch."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ntt-based-polynomial-multiplier-fpga_b200", "libnttb200.so")


def main(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = (0, 0, 0)
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and addr - tgt > best[0]:
                best = (addr - tgt, tgt, addr)
    _, lo, hi = best
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    mix = collections.Counter()
    for _, t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        mix[t.split()[0]] += 1
    print(f"# cuobjdump -sass libnttb200.so, kernel {fun}")
    print(f"# main loop, {len(body)} instructions, addresses {lo:#x}..{hi:#x} (whole kernel: {len(ins)})")
    print("# instruction mix of the loop body:")
    for op, c in mix.most_common():
        print(f"# {c:7d} {op}")
    heavy = sum(c * (2 if (".WIDE" in op or ".HI" in op) else 1) for op, c in mix.items() if op.startswith("IMAD"))
    print(f"# multiplier-pipe slots (IMAD 1, IMAD.WIDE/.HI 2): {heavy}")
    for a, t in body:
        print(f"/*{a:04x}*/  {t}")


if __name__ == "__main__":
    main(sys.argv[1])
