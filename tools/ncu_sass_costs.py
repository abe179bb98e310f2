#!/usr/bin/env python
"""Per-opcode and per-source-line ISSUE-COST table of one launch of an .ncu-rep (--import-source on, -lineinfo).

Cost model measured on B200 (profiles/r01_ubench.txt): an SM sub-partition spends 2 issue cycles on
the half-rate opcodes (packed FP32 x2, IMAD, LOP3, SHF, PRMT, ISETP, SEL, VIMNMX*, FMNMX, FSETP, FSEL, PLOP3, I2F/F2I, MUFU)
and 1 on the others (FADD/FMUL/FFMA scalar, IADD3, LEA, VIADD, memory, control).  cost = sum over executed warp
instructions.  usage: ncu_sass_costs.py report.ncu-rep launch_index [voxels]"""
import collections, csv, io, re, subprocess, sys

HALF = {"FADD2", "FMUL2", "FFMA2", "IMAD", "LOP3", "SHF", "PRMT", "ISETP", "SEL", "VIMNMX", "VIMNMX3", "FMNMX", "FMNMX3",
        "FSETP", "FSEL", "PLOP3", "I2F", "F2I", "MUFU", "IABS", "FLO", "BREV", "POPC", "I2FP", "F2FP", "VIADDMNMX", "IMNMX",
        "HFMA2", "HADD2", "HMUL2", "SGXT", "BMSK", "LOP", "FCHK", "DADD", "DMUL", "DFMA", "R2P", "P2R", "VABSDIFF", "VABSDIFF4"}


def opcode(s):
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", s)
    return m.group(1) if m else "?"


def main():
    rep, li = sys.argv[1], sys.argv[2]
    voxels = float(sys.argv[3]) if len(sys.argv) > 3 else 1024.0 ** 3
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", li,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # "sass,cuda" view: a row per CUDA line (Line No, Source, "-", "-", ...) followed by the SASS rows attributed to it
    # ("", "", Address, Source, ...)
    H, name, cur, cur_line = None, None, None, None
    ops, cost_by_op = collections.Counter(), collections.Counter()
    line_cost, line_ops, line_text = collections.Counter(), collections.defaultdict(collections.Counter), {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            name = r[1]
            continue
        if r and r[0] == "Line No":
            H = r
            ie = H.index("Instructions Executed")
            continue
        if H is None or len(r) < len(H):
            continue
        if r[0] != "":
            try:
                cur_line = (cur, int(r[0]))
                line_text[cur_line] = r[1]
            except ValueError:
                pass
            continue
        if not r[2].startswith("0x"):
            continue
        try:
            n = int(r[ie])
        except ValueError:
            continue
        op = opcode(r[3])
        c = n * (2 if op in HALF else 1)
        line_cost[cur_line] += c
        line_ops[cur_line][op] += n
    # opcode totals from the plain SASS view (the mixed view repeats an instruction under every frame of its inline stack)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", li,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    H = None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "Address":
            H = r
            ie = H.index("Instructions Executed")
            continue
        if H is None or len(r) < len(H) or not r[0].startswith("0x"):
            continue
        try:
            n = int(r[ie])
        except ValueError:
            continue
        op = opcode(r[1])
        ops[op] += n
        cost_by_op[op] += n * (2 if op in HALF else 1)
    # the source page's counters can be a multiple of the launch's real instruction count (several replay passes are
    # summed); normalise to smsp__inst_executed.sum of the raw page
    wv = voxels / 32.0
    rawp = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--launch-skip", li, "--launch-count", "1"],
                          capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(rawp)))
    try:
        true_inst = float(rr[2][rr[0].index("smsp__inst_executed.sum")].replace(",", ""))
        scale = sum(ops.values()) / true_inst
        wv *= scale
        print(f"(source-page counters = {scale:.2f} x smsp__inst_executed.sum; normalised)")
    except (ValueError, IndexError):
        pass
    tot_i, tot_c = sum(ops.values()), sum(cost_by_op.values())
    print(name)
    print(f"warp instr/32 voxels {tot_i / wv:.1f}   issue cost (2*half + full) {tot_c / wv:.1f} cycles per 32 voxels")
    print(f"{'opcode':10s} {'instr/vox':>9s} {'cost':>7s} {'cost%':>6s}")
    for op, c in sorted(cost_by_op.items(), key=lambda kv: -kv[1])[:28]:
        print(f"{op:10s} {ops[op] / wv:9.2f} {c / wv:7.2f} {c / tot_c * 100:5.1f}%{'  (half rate)' if op in HALF else ''}")
    tail(line_cost, line_ops, line_text, wv, tot_c)


def tail(line_cost, line_ops, line_text, wv, tot_c):
    print("\n(per line: an instruction is listed under every frame of its inline stack, so lines of callers and callees overlap)")
    print(f"{'file:line':26s} {'cost':>6s} {'cost%':>6s}  opcodes (instr per 32 voxels)   | source")
    for key, c in sorted(line_cost.items(), key=lambda kv: -kv[1])[:45]:
        f, ln = key
        mix = " ".join(f"{o}:{n / wv:.1f}" for o, n in line_ops[key].most_common(5))
        print(f"{f + ':' + str(ln):26s} {c / wv:6.1f} {c / tot_c * 100:5.1f}%  {mix:58s} | {line_text[key].strip()[:90]}")


if __name__ == "__main__":
    main()

