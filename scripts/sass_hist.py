#!/usr/bin/env python3
"""Static SASS opcode histogram per kernel of an object / cubin / .so (cuobjdump -sass): a quick instruction-mix view of the
pass kernels without a GPU.  usage: sass_hist.py <file> [function-substring] [--top N]"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, hist = None, {}
    pat = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in txt.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            hist[fn] = collections.Counter()
            continue
        m = pat.match(line)
        if m and fn:
            hist[fn][m.group(1).split(".")[0]] += 1
    for fn, h in hist.items():
        if filt and filt not in fn:
            continue
        tot = sum(h.values())
        print(f"== {fn}  total {tot}")
        print("   " + ", ".join(f"{k}:{v}" for k, v in h.most_common(top)))


if __name__ == "__main__":
    main()
