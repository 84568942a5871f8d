"""Instruction mix of a kernel's hottest loop from cuobjdump SASS (build-box aid: no GPU needed).

usage: python tools/sass_loop.py <object or .so> <substring of the mangled kernel name> [--dump]
Finds the backward branch with the longest body, and prints the per-opcode counts of that body, with the blocks that are
skipped by forward branches over >= 20 instructions (rare paths) counted separately.
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    body = None
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat in name:
            body = f
            print("kernel:", name)
            break
    if body is None:
        sys.exit("kernel not found")
    ins = []
    for l in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
    lo, hi = best
    print(f"loop 0x{lo:x}..0x{hi:x}: {addr[hi] - addr[lo] + 1} instructions")
    skip = []
    for i, (a, t) in enumerate(ins):
        if lo <= a <= hi:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt > a and tgt in addr and addr[tgt] - i >= 20 and tgt <= hi + 16:
                    skip.append((a + 16, tgt))
    main_c, rare_c = collections.Counter(), collections.Counter()
    for a, t in ins:
        if a < lo or a > hi:
            continue
        t = re.sub(r"^@!?U?P\d\s+", "", t)
        op = t.split()[0].split(".")[0]
        (rare_c if any(s <= a < e for s, e in skip) else main_c)[op] += 1
        if "--dump" in sys.argv:
            print(("    " if any(s <= a < e for s, e in skip) else "") + f"{a:05x} {t}")
    print("main path:", sum(main_c.values()), " skipped blocks:", sum(rare_c.values()), f"({len(skip)} blocks)")
    for k, v in main_c.most_common():
        print(f"  {k:10s} {v}")


if __name__ == "__main__":
    main()
