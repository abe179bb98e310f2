#!/usr/bin/env python
"""Static issue-cost per source line inside an address range of a kernel (nvdisasm --print-line-info; no GPU needed).

    tools/sass_line_cost.py obj.o 'kernel name fragment' 0xLO 0xHI [units]

Cost model as tools/sass_loop_cost.py (2 cycles for half-rate opcodes, 1 otherwise); the line is the innermost source
location nvdisasm reports for the instruction."""
import collections, os, re, subprocess, sys, tempfile

from sass_loop_cost import HALF


def main():
    obj, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    per = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    sec = None
    for s in re.split(r"\n\.text\.", txt)[1:]:
        if pat in s.split("\n")[0]:
            sec = s
            break
    cost, ops, src = collections.Counter(), collections.defaultdict(collections.Counter), {}
    cur = None
    for l in sec.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if m and lo <= int(m.group(1), 16) <= hi:
            op = m.group(3)
            cost[cur] += 2 if op in HALF else 1
            ops[cur][op] += 1
    tot = sum(cost.values())
    print(f"range 0x{lo:x}-0x{hi:x}: cost {tot} ({tot / per:.1f} per unit)")
    cache = {}
    for key, c in sorted(cost.items(), key=lambda kv: -kv[1])[:40]:
        f, ln = key if key else ("?", 0)
        text = ""
        for root in ("cuda_mesh_voxelization_b200/csrc", "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
            p = os.path.join(root, f)
            if os.path.exists(p):
                cache.setdefault(p, open(p, errors="replace").read().split("\n"))
                text = cache[p][ln - 1].strip() if ln - 1 < len(cache[p]) else ""
                break
        mix = " ".join(f"{o}:{n}" for o, n in ops[key].most_common(5))
        print(f"{f}:{ln:<5d} {c / per:6.1f}  {mix:52s} | {text[:100]}")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
