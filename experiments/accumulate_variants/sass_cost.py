#!/usr/bin/env python3
"""Static cost of a code path under the issue model measured on sm_100:
    cycles per warp and scheduler ~ 4 x IMAD.WIDE + 2 x (IMAD | IMAD.HI) + 1 x every other instruction.
(IMAD.HI at 2 cycles is what makes the production accumulate kernel come out at 5560 against 5530 measured; at 4 it would be 5714.)

    python scripts/sass_cost.py <lib.so|obj.o> <kernel name substring> <lo-hi>[,<lo-hi>...]

Address ranges (hex, half-open) select the hot path by hand from the kernel's branch structure, e.g. the accumulation loop
without its bucket-flush block and without the exceptional (P == Q) path.  Prints the census and the modelled cycles.
"""
import re
import signal
import subprocess
import sys
from collections import Counter


def kernel_rows(path, name):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    rows, on = [], False
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            on = name in m.group(1)
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            words = m.group(2).split()
            rows.append((int(m.group(1), 16), words[1] if words[0].startswith("@") else words[0]))
    return rows


def main():
    path, name, spec = sys.argv[1], sys.argv[2], sys.argv[3]
    ranges = [tuple(int(x, 16) for x in r.split("-")) for r in spec.split(",")]
    c = Counter(op for a, op in kernel_rows(path, name) if any(lo <= a < hi for lo, hi in ranges))
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
    narrow = c.get("IMAD", 0) + c.get("IMAD.HI.U32", 0)
    total = sum(c.values())
    other = total - wide - narrow
    print(f"{name}: {total} instructions: {wide} IMAD.WIDE, {narrow} IMAD/IMAD.HI, {other} other -> {4 * wide + 2 * narrow + other} cycles")
    for k, v in c.most_common(16):
        print(f"    {k:24s}{v:6d}")


if __name__ == "__main__":
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # "| head" is a normal way to use this
    main()
