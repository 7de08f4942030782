#!/usr/bin/env python
"""Static instruction mix of the GGS kernel's packed fast path (bulk-async ring walk), plain vs paired stream layout.

    python tools/ggs_sass_stats.py            # needs build/obj/api_core.o (python -m posediffusion_b200.build), no GPU

The fast path of one 4-round chunk is the straight-line block between the ring loads (`LDS.128`) that follow the mbarrier
wait and the branch that skips the slow path.  The script finds the block with the most FFMA2 instructions in each
non-eval instantiation and prints its instruction histogram."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "obj", "api_core.o")


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, res = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            res[name] = []
        elif name and re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", line):
            ins = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip()
            res[name].append(ins.split("*/", 1)[1].strip())
    return res


def basic_blocks(ins):
    block = []
    for i in ins:
        block.append(i)
        op = i.split()[1] if i.startswith("@") else i.split()[0]
        if op.startswith(("BRA", "BSYNC", "EXIT", "RET", "WARPSYNC", "BSSY", "CALL")):
            yield block
            block = []
    if block:
        yield block


def opcode(i):
    tok = i.split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    return op.split(".")[0].rstrip(";")


def main():
    funcs = functions(OBJ)
    rows = []
    for tag, suffix in (("plain", "ggs_entryILb0ELb0EEE"), ("paired", "ggs_entryILb0ELb1EEE")):
        name = next(n for n in funcs if suffix in n)
        best = max(basic_blocks(funcs[name]), key=lambda b: sum(opcode(i) == "FFMA2" for i in b))
        hist = collections.Counter(opcode(i) for i in best)
        rows.append((tag, len(funcs[name]), len(best), hist))
    for tag, total, n, hist in rows:
        arith = sum(hist[k] for k in ("FFMA2", "FMUL2", "FADD2", "MUFU", "FSETP", "FSEL", "FSET"))
        moves = hist["MOV"] + hist["IMAD"]  # IMAD.MOV.U32 is a register move too
        print(f"{tag:7s}: kernel {total} instructions; fast-path block of one 4-round chunk (128 matches per warp): {n} instructions "
              f"= {n / 4:.1f} per 32-match round; arithmetic {arith}, register moves {moves}, LDS {hist['LDS']}")
        print("         " + ", ".join(f"{k} {v}" for k, v in hist.most_common()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
