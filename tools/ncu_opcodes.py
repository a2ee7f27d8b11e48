"""Aggregate executed warp-instructions by opcode from an ncu report's source page (needs -lineinfo / --import-source)."""
import collections
import csv
import subprocess
import sys


def main(path, units=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    tot = collections.Counter()
    stall = collections.Counter()
    for r in rows:
        if len(r) > 6 and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < 8 or not r[0].startswith("0x"):
            continue
        try:
            n = int(float(r[hdr.index("Instructions Executed")]))
        except ValueError:
            continue
        toks = r[1].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDG", "LDS", "STS")) else op.split(".")[0]
        tot[op] += n
        stall[op] += int(float(r[hdr.index("# Samples")] or 0))
    s = sum(tot.values())
    print(f"total warp-instructions {s}" + (f"  = {s / units:.1f} per unit" if units else ""))
    tot_samples = sum(stall.values()) or 1
    for op, n in tot.most_common(28):
        extra = f" {n / units:7.2f}/unit" if units else ""
        print(f"  {op:14s} {n:11d} {100 * n / s:5.1f}%{extra}   stall-samples {100 * stall[op] / tot_samples:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
