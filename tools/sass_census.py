"""SASS opcode census of every kernel in libfp4_b200's objects (cuobjdump -sass), written to profiles/.  The lines
that matter for "is this Blackwell-native": UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA),
IMMA (mma.sync u8 x s8, the GEMV's integer tensor-core path), LDGSTS (cp.async)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "torch_bnb_fp4_b200", "build")
KEY = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "IMMA", "HMMA", "LDGSTS",
       "SYNCS", "PRMT", "LOP3", "IMAD", "FFMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ACQBULK", "PREEXIT")


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:  # noqa: BLE001
        return n


def main(out_path):
    lines = []
    for obj in sorted(f for f in os.listdir(BUILD) if f.endswith(".o") and f.count(".") == 1):
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        cur, cnt = None, collections.OrderedDict()
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = m.group(1)
                cnt[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m and cur:
                cnt[cur][m.group(1).split(".")[0]] += 1
        for fn, c in cnt.items():
            name = re.sub(r"fp4b200::\(anonymous namespace\)::", "", demangle(fn))
            name = re.sub(r"\(fp4b200.*", "", name)[:110]
            total = sum(c.values())
            keys = "  ".join(f"{k}={c[k]}" for k in KEY if c.get(k))
            lines.append(f"{obj:18s} {name:110s} total={total:5d}  {keys}")
    with open(out_path, "w") as f:
        f.write("SASS opcode census per kernel (static instruction counts; cuobjdump -sass of torch_bnb_fp4_b200/build/*.o, "
                "sm_100a)\n" + "\n".join(lines) + "\n")
    print(f"{len(lines)} kernels -> {out_path}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_census.txt"))
