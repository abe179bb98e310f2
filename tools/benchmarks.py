#!/usr/bin/env python
"""Benchmark runner with the command line, label parsing and CSV layout of the reference's scripts/benchmarks.py
(SURVEY §8 f3), so its plotting scripts (scripts/plot_*.py, generate_plots.py) read our CSVs unchanged.

For every mesh file in --folder, every back-end in --types and every power-of-two grid side in [--minsize, --maxsize] it
runs   cli <mesh> -n<size> -t<type> -m<niter> -p1 [-s]   (-s up to 512^3 unless --no-sdf, like the reference), parses the
"[Label(...)::Sub]: X ms" lines the CLI prints, and writes  <output>/<mesh>/<mesh>_<main_label>.csv  with one row per
iteration: size, then the main label's total and its ::memory / ::processing parts in snake case.

Differences from the reference script: the executable defaults to this repository's CLI (apps/cli/cli, override with
--exec) and --types defaults to 4 (the B200 back-end; the reference's own CLI knows 0-3).
"""
from __future__ import annotations

import argparse
import csv
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LABEL = re.compile(r"\[(.*)\]: ([\d.]+) ms")


def snake(label: str) -> str:
    """B200Vox::Memory -> b200_vox__memory (scope separator becomes a double underscore)."""
    out = label.replace("::", "__")
    out = re.sub(r"(?<=[a-z0-9])([A-Z])", r"_\1", out)
    out = re.sub(r"([A-Z]+)([A-Z][a-z])", r"\1_\2", out)
    return re.sub(r"__+", "__", out.lower())


def parse_run(stdout: str, size: str, table: dict) -> None:
    """Sub-scope lines (Label::Sub) precede their main line (Label): they are summed into the record the main line closes."""
    pending: dict = {}
    for line in stdout.splitlines():
        m = LABEL.search(line)
        if not m:
            continue
        label = re.sub(r"\s*\(.*?\)", "", m.group(1))          # drop "(mesh name)"
        full, main = snake(label), snake(label.split("::")[0])
        pending[full] = pending.get(full, 0.0) + float(m.group(2))
        if "__" not in full:                                     # the main line: one iteration is complete
            table[main][size].append(dict(pending))
            pending.clear()


def main() -> int:
    ap = argparse.ArgumentParser(description="Benchmark runner (reference-compatible CSVs)")
    ap.add_argument("--niter", type=int, default=10, help="Number of iterations per test")
    ap.add_argument("--folder", type=str, default="./tests", help="Folder with input files")
    ap.add_argument("--maxsize", type=int, default=128, help="Maximum size (power of 2, starting from --minsize)")
    ap.add_argument("--minsize", type=int, default=32, help="Minimum size")
    ap.add_argument("--output", type=str, default="benchmarks", help="Output folder for CSVs")
    ap.add_argument("--no-sdf", action="store_true", help="Never pass -s")
    ap.add_argument("--types", nargs="+", default=["4"], help="Back-ends to run (4 = b200)")
    ap.add_argument("--exec", dest="exe", default=str(ROOT / "apps" / "cli" / "cli"), help="CLI executable")
    args = ap.parse_args()

    sizes, s = [], args.minsize
    while s <= args.maxsize:
        sizes.append(str(s))
        s *= 2
    out_root = Path(args.output)
    out_root.mkdir(exist_ok=True)
    for mesh in sorted(Path(args.folder).iterdir()):
        if not mesh.is_file() or mesh.suffix.lower() != ".obj":
            continue
        table: dict = defaultdict(lambda: defaultdict(list))     # main label -> size -> [ {column: ms} per iteration ]
        for typ in args.types:
            for size in sizes:
                cmd = [args.exe, str(mesh), f"-n{size}", f"-t{typ}", f"-m{args.niter}", "-p1"]
                if not args.no_sdf and int(size) <= 512:
                    cmd.append("-s")
                print("Running:", " ".join(cmd), flush=True)
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                if r.returncode != 0:
                    print(f"{args.exe} failed on {mesh.name} (exit {r.returncode})\nSTDOUT:\n{r.stdout}\nSTDERR:\n{r.stderr}", file=sys.stderr)
                    return 1
                parse_run(r.stdout, size, table)
        folder = out_root / mesh.stem
        folder.mkdir(exist_ok=True)
        for main_label, by_size in table.items():
            columns = sorted({c for rows in by_size.values() for row in rows for c in row})
            with open(folder / f"{mesh.stem}_{main_label}.csv", "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["size"] + columns)
                for size in sorted(by_size, key=int):
                    for row in by_size[size]:
                        w.writerow([size] + [row.get(c, "") for c in columns])
    return 0


if __name__ == "__main__":
    sys.exit(main())
