#!/usr/bin/env python
"""Static pipe census of the fine-pass traversal loop from cuobjdump SASS (which issue pipe each
instruction of the loop body occupies). The loop is ALU-pipe bound."""
import re
import subprocess
import sys

FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA2", "FMUL2", "FADD2")
ALU = ("IADD3", "LOP3", "SHF", "FMNMX", "FMNMX3", "FSEL", "SEL", "ISETP", "FSETP", "MOV", "PRMT", "VIADD", "LEA", "R2P",
       "P2R", "POPC", "FLO", "IABS", "BREV", "PLOP3", "VIMNMX", "CS2R", "FCHK")
LSU = ("LDG", "STG", "LDS", "STS", "LDC", "LDCU", "ATOM", "RED", "LDL", "STL")
CTL = ("BRA", "BSSY", "BSYNC", "BREAK", "EXIT", "CALL", "RET", "WARPSYNC", "NOP", "BAR")


def pipe(op):
    base = op.split(".")[0]
    for name, group in (("fma", FMA), ("alu", ALU), ("lsu", LSU), ("ctl", CTL)):
        if base in group:
            return name
    return "other"


def main():
    obj = sys.argv[1] if len(sys.argv) > 1 else "sparse-voxel-octrees_b200/build/svo_kernels.o"
    pat = sys.argv[2] if len(sys.argv) > 2 else "finePassKernelILb1EjEE"
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    on = False
    lines = []
    for ln in sass.splitlines():
        if "Function :" in ln:
            on = pat in ln
        elif on:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", ln)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
    # loop = from the ISETP that guards the descriptor fetch to the backward BRA
    start = next(i for i, (_, t) in enumerate(lines) if t.startswith("ISETP.NE.AND") and "RZ" in t and
                 any("LDG" in lines[j][1] for j in range(i, min(i + 6, len(lines)))))
    head_addr = lines[start][0]
    end = max(i for i, (_, t) in enumerate(lines) if re.search(r"BRA\s+0x%x\b" % head_addr, t))
    census = {}
    for _, t in lines[start:end + 1]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        p = pipe(t.split()[0])
        census[p] = census.get(p, 0) + 1
    print(f"loop body {end - start + 1} instructions (all paths, shading included):", census)


if __name__ == "__main__":
    main()
