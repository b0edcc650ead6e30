#!/usr/bin/env python3
"""Emit one translation unit per FFT length (gen/len_<N>.cu) plus gen/registry.cpp.

A plan is (N, R1, R2, R3, T, W): N = R1*R2*R3 (R3 == 1 -> two shared-memory stages), T threads cooperate
on one line, W lines per CTA.  W = 16 (128-byte global segments) while the W x N complex tile stays <= ~74 KB so
that three CTAs fit in one SM's 227 KB of shared memory; W = 8 above that.
"""
import os
import sys

PLANS = [
    # N,   R1, R2, R3, T,  W
    (32, 8, 4, 1, 8, 16), (36, 6, 6, 1, 6, 16), (40, 8, 5, 1, 8, 16), (48, 8, 6, 1, 8, 16),
    (50, 10, 5, 1, 10, 16), (54, 9, 6, 1, 9, 16), (60, 10, 6, 1, 10, 16), (64, 8, 8, 1, 8, 16),
    (72, 9, 8, 1, 9, 16), (80, 10, 8, 1, 10, 16), (90, 10, 9, 1, 10, 16), (96, 12, 8, 1, 12, 16),
    (100, 10, 10, 1, 10, 16), (108, 12, 9, 1, 12, 16), (120, 12, 10, 1, 12, 16), (128, 16, 8, 1, 16, 16),
    (144, 12, 12, 1, 12, 16), (150, 15, 10, 1, 15, 16), (160, 16, 10, 1, 16, 16), (180, 15, 12, 1, 15, 16),
    (192, 16, 12, 1, 16, 16), (200, 20, 10, 1, 20, 16), (216, 18, 12, 1, 18, 16), (240, 16, 15, 1, 16, 16),
    (256, 16, 16, 1, 16, 16), (270, 18, 15, 1, 18, 16), (288, 18, 16, 1, 18, 16), (300, 20, 15, 1, 20, 16),
    (320, 20, 16, 1, 20, 16), (360, 20, 18, 1, 20, 16), (384, 8, 8, 6, 16, 16), (400, 20, 20, 1, 20, 16),
    (432, 12, 6, 6, 18, 16), (450, 15, 6, 5, 15, 16), (480, 10, 8, 6, 16, 16), (500, 10, 10, 5, 25, 16),
    (512, 8, 8, 8, 32, 16), (540, 15, 6, 6, 18, 16), (576, 12, 12, 4, 24, 16),
    (600, 10, 10, 6, 20, 8), (640, 10, 8, 8, 16, 8), (720, 12, 10, 6, 12, 8), (768, 8, 8, 12, 32, 8),
    (800, 10, 10, 8, 20, 8), (864, 12, 12, 6, 24, 8), (900, 15, 10, 6, 30, 8), (960, 10, 12, 8, 24, 8),
    (1000, 10, 10, 10, 25, 8), (1024, 16, 8, 8, 32, 8), (1080, 15, 12, 6, 18, 8), (1152, 12, 12, 8, 24, 8),
]


# hand-picked x plans "r1,r2,r3[,xt]" (measured on B200); everything else is chosen by x_plan()
XPLAN_OVERRIDE = {
    540: (18, 30, 1, 32),                       # measured against (15,6,6) and (20,27)
    # three-stage lengths: last-stage radix 6 or 10 (no padding inside a line -> bulk-copy staging, conflict-free as is)
    576: (12, 8, 6), 640: (8, 8, 10), 768: (16, 8, 6), 800: (10, 8, 10), 960: (12, 8, 10), 1152: (12, 16, 6),
}


def col_plan(n, r1, r2, r3, t, w):
    """Plan of the column (y / z) kernels.  Lengths from 384 to 640 that factor as a * b with a <= 32 and b <= 20 run as TWO stages,
    larger radix first, one first-stage butterfly per thread (T = b threads per column, <= 320 threads per CTA): 4 instead of 6
    (convolution: 7 instead of 11) shared-memory / global accesses per element, 3 instead of 5 barrier phases, and up to 32
    independent global loads in flight per thread.  Needs ~96-118 registers -> two CTAs per SM.  Measured on c3 (N = 540): y forward
    0.47 -> 0.37 ms, z convolution 0.94 -> 0.85 ms against (15, 6, 6); (27, 20) and (30, 18) are equivalent for y, (27, 20) is 2 %
    faster in the convolution.  b > 20 (e.g. 24 x 24) would need 384 threads, i.e. <= 85 registers: the convolution stage spills."""
    forced = os.environ.get(f"MVD_CPLAN_{n}")                                   # "r1,r2,r3,t,w"
    if forced:
        return tuple(int(x) for x in forced.split(","))
    if n < 384 or n > 640 or r3 == 1:
        return r1, r2, r3, t, w
    cands = []
    for b in range(8, 21):
        if n % b or not radix_ok(b):
            continue
        a = n // b
        if a < b or a > 32 or not radix_ok(a):
            continue
        cands.append((a, b))
    if not cands:
        return r1, r2, r3, t, w
    a, b = min(cands)
    return a, b, 1, b, 16


def radix_ok(r):
    for f in (2, 3, 5):
        while r % f == 0:
            r //= f
    return r == 1


def x_plan(n, r1, r2, r3):
    """Plan of the x kernels.  Two shared-memory stages whenever N = a*b with a <= 20 (the stage whose butterfly stays in registers
    across the real-space work) and b <= 30: half the shared-memory traffic and barriers of a three-stage plan (measured on c3:
    quotient pass -9 %).  The last-stage radix b avoids multiples of 4 when it can (no padding inside a line: the lines are staged
    by bulk copies in the persistent kernels, and odd or 2-mod-4 radices are bank-conflict free as they are); among those the
    most balanced pair wins.  XT threads per line: N / r1 butterflies, rounded up to a whole warp when that wastes <= 10 % of the
    lanes (a line never straddles a warp then).  XL lines per CTA: ~256 threads, line tile <= 36 KB (two tiles + tables per
    persistent CTA, two CTAs per SM)."""
    forced = os.environ.get(f"MVD_XPLAN_{n}") or XPLAN_OVERRIDE.get(n)          # "r1,r2,r3[,xt]"
    xt_forced = None
    if forced:
        f = [int(x) for x in (forced.split(",") if isinstance(forced, str) else forced)]
        xr1, xr2, xr3 = f[0], f[1], f[2]
        assert xr1 * xr2 * xr3 == n
        xt_forced = f[3] if len(f) > 3 else None
    else:
        cands = []
        for a in range(2, 21):
            if n % a or not radix_ok(a):
                continue
            b = n // a
            if b < 4 or a < 4 or b > 30 or not radix_ok(b):
                continue
            pad = 1 if b % 4 == 0 else 0
            vec = 0 if b % 2 == 0 else 1
            cands.append((pad, max(a, b), vec, a, b))
        if cands:
            _, _, _, xr1, xr2 = min(cands)
            xr3 = 1
        else:
            xr1, xr2, xr3 = r1, r2, r3
    rl = xr3 if xr3 > 1 else xr2
    pad = 2 if rl % 4 == 0 else 0
    ls = n + pad * (n // rl)
    xt = n // xr1
    if xt % 32 and (-xt) % 32 <= 0.10 * xt:
        xt += (-xt) % 32
    xt = xt_forced or xt
    want = 256
    choice = None
    for xl in range(1, 129):
        thr = xt * xl
        smem = xl * ls * 8
        if thr > 512 or smem > 36 * 1024:
            break
        score = abs(thr - want) + (0 if thr % 32 == 0 else 40)
        if choice is None or score < choice[0]:
            choice = (score, xl)
    xl = choice[1] if choice else 1
    if os.environ.get(f"MVD_XL_{n}"):
        xl = int(os.environ[f"MVD_XL_{n}"])
    # line groups of the persistent (bulk-copy staged) kernels: half the group of the one-shot kernels while that leaves >= 128 threads
    # (more resident CTAs = more independent barrier domains per SM; measured on c3: forward -5 %, quotient -7 %)
    xlp = xl // 2 if (xl % 2 == 0 and xt * (xl // 2) >= 128) else xl
    if os.environ.get(f"MVD_XLP_{n}"):
        xlp = int(os.environ[f"MVD_XLP_{n}"])
    return xr1, xr2, xr3, xt, xl, xlp


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "gen")
    only = os.environ.get("MVD_LENGTHS")
    plans = PLANS
    if only:
        keep = {int(x) for x in only.replace(",", " ").split()}
        plans = [p for p in PLANS if p[0] in keep]
    os.makedirs(out, exist_ok=True)
    wanted = set()

    def emit(path, text):
        wanted.add(os.path.basename(path))
        if os.path.exists(path) and open(path).read() == text:
            return
        with open(path, "w") as f:
            f.write(text)

    for (n, r1, r2, r3, t, w) in plans:
        assert r1 * r2 * r3 == n, n
        xr1, xr2, xr3, xt, xl, xlp = x_plan(n, r1, r2, r3)
        r1, r2, r3, t, w = col_plan(n, r1, r2, r3, t, w)
        emit(os.path.join(out, f"len_{n}.cu"),
             f'// generated by gen_lengths.py -- do not edit\n#include "len_ops_impl.cuh"\n'
             f'MVD_DEFINE_LEN({n}, {r1}, {r2}, {r3}, {t}, {w}, {xr1}, {xr2}, {xr3}, {xt}, {xl}, {xlp})\n')
    decl = "".join(f"const LenOps* len_ops_{p[0]}();\n" for p in plans)
    table = ",\n    ".join(f"len_ops_{p[0]}()" for p in plans)
    lens = ", ".join(str(p[0]) for p in plans)
    emit(os.path.join(out, "registry.cpp"),
         "// generated by gen_lengths.py -- do not edit\n#include \"backend.h\"\nnamespace mvd {\n" + decl +
         "const LenOps* find_len_ops(int N) {\n    static const LenOps* const all[] = {\n    " + table +
         "};\n    for (const LenOps* o : all) if (o->N == N) return o;\n    return nullptr;\n}\n"
         "const std::vector<int>& supported_lengths() {\n    static const std::vector<int> v = {" + lens + "};\n    return v;\n}\n}\n")
    # drop stale generated files
    for fn in os.listdir(out):
        if (fn.startswith("len_") and fn.endswith(".cu") or fn == "registry.cpp") and fn not in wanted:
            os.remove(os.path.join(out, fn))
    print(" ".join(str(p[0]) for p in plans))


if __name__ == "__main__":
    main()
