"""Static evidence from the built library (no GPU needed): per kernel, the SASS mnemonics that prove which hardware path it uses
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = bulk copy, HMMA = mma.sync, LDSM = ldmatrix, SYNCS =
mbarrier), plus registers / spills from the ptxas log.  `python tools/sass_summary.py > profiles/rNN_sass_summary.md`."""
from __future__ import annotations

import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastllm_b200", "libfastllm_b200.so")
PTXAS_LOG = os.path.join(ROOT, "fastllm_b200", "_obj", "ptxas.log")
MNEMONICS = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG.2D", "UTMALDG.3D", "UBLKCP", "HMMA.16816.F32.BF16", "LDSM", "SYNCS", "STL", "LDL"]
_CUDA_BIN = "/usr/local/cuda/bin"


def _tool(name: str) -> str:
    p = os.path.join(_CUDA_BIN, name)
    return p if os.path.exists(p) else name


def demangle(names):
    out = subprocess.run([_tool("cu++filt")] + list(names), capture_output=True, text=True)
    if out.returncode != 0 or not out.stdout.strip():
        out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True, check=True)
    return [re.sub(r"\(.*", "", line.strip().replace("(int)", "")).replace("void ", "").replace("fl::", "") for line in out.stdout.strip().splitlines()]


def kernel_mnemonics(lib: str = LIB) -> dict:
    """mangled kernel name -> Counter(mnemonic prefix -> occurrences)."""
    sass = subprocess.run([_tool("cuobjdump"), "-sass", lib], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = res.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            for k in MNEMONICS:
                if op == k or op.startswith(k + ".") or (k.endswith("LDSM") and op.startswith("LDSM")):
                    cur[k] += 1
    return res


def ptxas_resources(log: str = PTXAS_LOG) -> dict:
    """mangled kernel name -> (registers, stack bytes, spill-store bytes, spill-load bytes)."""
    txt = open(log).read()
    res = {}
    for m in re.finditer(r"Function properties for (\S+)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s*: Used (\d+) registers", txt):
        res[m.group(1)] = (int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)))
    return res


def main():
    mn, rs = kernel_mnemonics(), ptxas_resources()
    names = sorted(mn)
    pretty = dict(zip(names, demangle(names)))
    cols = ["UTCHMMA", "LDTM", "UTMALDG.2D", "UTMALDG.3D", "UBLKCP", "HMMA.16816.F32.BF16", "LDSM", "SYNCS"]
    print("| kernel | regs | spill st/ld (B) | " + " | ".join(c.split(".")[0] if c.startswith("HMMA") else c for c in cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    for n in sorted(names, key=lambda k: pretty[k]):
        r = rs.get(n, (0, 0, 0, 0))
        if not any(mn[n][c] for c in cols) and r[2] == 0:
            continue                                # plain element-wise kernels: nothing to show
        print(f"| `{pretty[n][:80]}` | {r[0]} | {r[2]}/{r[3]} | " + " | ".join(str(mn[n][c] or "") for c in cols) + " |")


if __name__ == "__main__":
    main()
