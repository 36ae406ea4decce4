#!/usr/bin/env python
"""Static look at a kernel's SASS (no GPU needed): decodes the scheduling control word of every instruction (stall count,
yield, scoreboards, operand-reuse flags) and simulates the operand-reuse cache to count how many matrix-form FFMA2
(scalar x pair + pair: five distinct source registers) get an operand from it.
   cuobjdump -sass rgbd_pose_estimation_b200/csrc/_obj/score.o > /tmp/all.txt
   python tools/sass_control.py /tmp/all.txt score3d_raw_kernelILi2ELi1024ELi512ELi1ELi4ELi1 [--dump 2a00 3400]
Bit positions of the 128-bit instruction word (Volta and later): stall 105-108, yield 109, write barrier 110-112,
read barrier 113-115, wait mask 116-121, reuse 122-125."""
import re
import sys


def kernel_lines(path, name):
    out, on = [], False
    for l in open(path):
        if "Function : " in l:
            on = name in l
        if on:
            out.append(l.rstrip("\n"))
    return out


def decode(lines):
    ins, i = [], 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
                ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=(w >> 105) & 0xF, hold=(w >> 109) & 1,
                                wbar=(w >> 110) & 7, rbar=(w >> 113) & 7, wait=(w >> 116) & 0x3F, reuse=(w >> 122) & 0xF))
                i += 2
                continue
        i += 1
    return ins


def reuse_stats(ins):
    cache, tot, hit = {}, 0, 0
    for d in ins:
        parts = d["text"].split(None, 1)
        op = parts[0]
        if op.startswith("@") and len(parts) > 1:
            parts = parts[1].split(None, 1)
            op = parts[0]
        ops = [o.strip() for o in parts[1].split(",")] if len(parts) > 1 else []
        if op.startswith(("BRA", "BSYNC", "BAR")):
            cache = {}
        mat = op == "FFMA2" and len(ops) == 4 and ops[1].endswith(".F32") and not ops[1].startswith("-")
        anyhit = False
        for si, o in enumerate(ops[1:]):
            reg = o.replace(".reuse", "").lstrip("-|").split(".")[0].rstrip("|")
            if not reg.startswith("R"):
                continue
            if cache.get(si) == reg:
                anyhit = True
            if ".reuse" in o:
                cache[si] = reg
        if mat:
            tot += 1
            hit += anyhit
    return tot, hit


if __name__ == "__main__":
    lines = kernel_lines(sys.argv[1], sys.argv[2])
    ins = decode(lines)
    tot, hit = reuse_stats(ins)
    print(f"{len(ins)} instructions; matrix FFMA2: {tot}, with an operand from the reuse cache: {hit}")
    if "--dump" in sys.argv:
        k = sys.argv.index("--dump")
        lo, hi = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
        for d in ins:
            if lo <= d["addr"] <= hi:
                print(f"{d['addr']:04x} stall={d['stall']} hold={d['hold']} wbar={d['wbar']} rbar={d['rbar']} wait={d['wait']:02x} "
                      f"reuse={d['reuse']:x}  {d['text'][:100]}")
