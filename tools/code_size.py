"""Static code size of a kernel by source function/line range (nvdisasm -g line info). Development aid."""
import re, subprocess, sys, os, tempfile, collections
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
lib = str(ROOT / "acmmp-spherical_b200/lib/libacmmp_b200.so")
kern = sys.argv[1] if len(sys.argv) > 1 else "k_passILi0E"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sec = None; cur = None; cnt = collections.Counter(); total = 0
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m: sec = m.group(1); continue
    if sec is None or kern not in sec: continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/", l): cnt[cur] += 1; total += 1
src = {}
def lines_of(f):
    if f not in src:
        p = ROOT / "acmmp-spherical_b200/csrc" / f
        src[f] = p.read_text().split("\n") if p.exists() else []
    return src[f]
def func_of(f, n):
    L = lines_of(f)
    for i in range(min(n, len(L)) - 1, -1, -1):
        l = L[i]
        if l and not l[0].isspace() and not l.startswith(("//", "#", "}", "template", "static_assert", "constexpr")) and "(" in l:
            m = re.search(r"(\w+)\s*\(", l)
            return m.group(1) if m else l[:30]
    return "?"
byf = collections.Counter()
for (k, c) in cnt.items():
    if k is None: byf["<none>"] += c; continue
    f, n = k
    key = f + ":" + func_of(f, n) if f.endswith((".cuh", ".cu")) else f
    byf[key] += c
print("total", total)
for k, c in byf.most_common(30): print(f"{c:6d}  {c/total*100:5.1f}%  {k}")
