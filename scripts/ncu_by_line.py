"""Executed warp-instructions and stall samples per CUDA source line, by joining the ncu SASS
page with nvdisasm's line info of the SAME build of libbatchdrones.so.

    python scripts/ncu_by_line.py <report.ncu-rep> <mangled-kernel-substring> [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "marl_gym_pybullet_drones_b200", "libbatchdrones.so")


def sass_lines(kernel_sub):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, capture_output=True)
        txt = ""
        for cub in sorted(os.listdir(td)):   # one cubin per translation unit; take the one with the kernel
            t = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout
            if any(l.strip().startswith(".section") and ".text." in l and kernel_sub in l for l in t.split("\n")):
                txt = t
                break
    lines = txt.split("\n")
    start = next(i for i, l in enumerate(lines)
                 if l.strip().startswith(".section") and ".text." in l and kernel_sub in l)
    out, cur = [], None
    for l in lines[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main(rep, kernel_sub, top=40):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    data = [r for r in rows[2:] if len(r) >= 6 and r[0].startswith("0x")]
    first, seen = [], set()
    for r in data:
        if r[0] in seen:
            break
        seen.add(r[0])
        first.append(r)
    sl = sass_lines(kernel_sub)
    if len(sl) != len(first):
        raise SystemExit(f"build mismatch: {len(sl)} SASS instructions in the library vs {len(first)} in the report")
    ex, st = defaultdict(int), defaultdict(int)
    for (loc, _), r in zip(sl, first):
        ex[loc] += int(r[5])
        st[loc] += int(r[2])
    tot, tots = sum(ex.values()), sum(st.values())
    print(f"total executed warp-instructions {tot}, stall samples {tots}")
    cache = {}
    order = (lambda kv: -st[kv[0]]) if os.environ.get("BY_STALL") else (lambda kv: -kv[1])
    for loc, n in sorted(ex.items(), key=order)[:top]:
        text = ""
        if loc:
            f = os.path.join(ROOT, "marl_gym_pybullet_drones_b200", "csrc", loc[0])
            if os.path.exists(f):
                cache.setdefault(f, open(f).read().split("\n"))
                text = cache[f][loc[1] - 1].strip()[:80]
        print(f"{100 * n / tot:5.1f}% inst {100 * st[loc] / max(tots, 1):5.1f}% stall  {loc}  {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
