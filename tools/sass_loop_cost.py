#!/usr/bin/env python
"""Static issue-cost of the loops of a kernel, from `cuobjdump -sass` text (no GPU needed).

    tools/sass_loop_cost.py build/jfa_flood4.o 'jfa_pass_flood4ILi2ELi16ELb0' [voxels_per_thread_per_iteration]

Cost model (B200): 2 issue cycles for the half-rate opcodes, 1 for the others.  A loop is an
address range [target, branch] of a backward branch; nested ranges are listed separately (the outer one includes the
inner).  Conditional blocks inside a loop are counted as if always executed, so compare variants, not absolutes."""
import collections, re, subprocess, sys

HALF = {"FADD2", "FMUL2", "FFMA2", "IMAD", "LOP3", "SHF", "PRMT", "ISETP", "SEL", "VIMNMX", "VIMNMX3", "FMNMX", "FMNMX3",
        "FSETP", "FSEL", "PLOP3", "I2F", "F2I", "MUFU", "IABS", "FLO", "BREV", "POPC", "I2FP", "F2FP", "VIADDMNMX", "IMNMX",
        "HFMA2", "HADD2", "HMUL2", "SGXT", "BMSK", "LOP", "FCHK", "R2P", "P2R"}


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0]
        if pat not in name:
            continue
        ins = []
        for l in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)(\S*)\s*(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), m.group(4), m.group(5)))
        print(name[-70:], f"{len(ins)} instructions")
        loops = []
        for a, op, mod, rest in ins:
            if op == "BRA":
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t and int(t.group(1), 16) <= a:
                    loops.append((int(t.group(1), 16), a))
        for lo, hi in sorted(loops, key=lambda x: x[0] - x[1]):
            body = [(op) for a, op, mod, rest in ins if lo <= a <= hi]
            c = collections.Counter(body)
            cost = sum(n * (2 if op in HALF else 1) for op, n in c.items())
            print(f"  loop 0x{lo:04x}-0x{hi:04x}: {len(body)} instr, cost {cost} ({cost / per:.1f} per unit)  " +
                  " ".join(f"{o}:{n}" for o, n in c.most_common(14)))


if __name__ == "__main__":
    main()
