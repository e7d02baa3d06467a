#!/usr/bin/env python3
"""Which kernels of two builds differ?  Compares `cuobjdump -sass` of two .so / .o files function by function,
instructions AND encodings (the encodings carry the scheduling control bits), whitespace-normalised.

    python scripts/sass_diff.py old.so new.so

Used before shipping a library whose GPU tests cannot be re-run: a change that only ADDS kernels must leave
every existing kernel byte-identical.
"""
import re
import signal
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        if line.startswith("Fatbin elf code"):
            cur = None
        line = " ".join(line.split())
        if cur and line:
            d[cur].append(line)
    return d


def main():
    a, b = functions(sys.argv[1]), functions(sys.argv[2])
    missing = [k for k in a if k not in b]
    changed = [k for k in a if k in b and a[k] != b[k]]
    new = [k for k in b if k not in a]
    print(f"{len(a)} kernels before, {len(b)} after")
    for title, names in (("removed", missing), ("changed", changed), ("added", new)):
        print(f"{title}: {len(names)}")
        for k in names:
            print("   ", k)
    sys.exit(1 if (missing or changed) else 0)


if __name__ == "__main__":
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # "| head" is a normal way to use this
    main()
