"""Attribute executed instructions of a profiled kernel (ncu --page source --csv --print-source sass) to
source lines, using nvdisasm -g line info of the SAME build.  Development aid.
usage: prof_lines.py <sass_csv> <kernel-substring> [lib.so]"""
import csv, re, subprocess, sys, collections, tempfile, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
csvf, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else str(ROOT / "acmmp-spherical_b200/lib/libacmmp_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sec = None; cur = None; lines = {}
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m: sec = m.group(1); continue
    if sec is None or kern not in sec: continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m: lines[int(m.group(1), 16)] = cur
rows = []; hdr = None; seen = set(); base = None
for line in csv.reader(open(csvf)):
    if line and line[0] == "Address": hdr = line; continue
    if hdr and len(line) == len(hdr):
        d = dict(zip(hdr, line))
        if d["Address"] in seen: break
        seen.add(d["Address"]); rows.append(d)
base = int(rows[0]["Address"], 16)
agg = collections.Counter(); samp = collections.Counter()
tot = 0
for d in rows:
    off = int(d["Address"], 16) - base
    n = int(d["Instructions Executed"]); tot += n
    agg[lines.get(off)] += n; samp[lines.get(off)] += int(d["# Samples"])
src = {}
def text(f, n):
    if f not in src:
        p = ROOT / "acmmp-spherical_b200/csrc" / f
        src[f] = p.read_text().split("\n") if p.exists() else []
    return src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
print("total", tot)
for (k, n) in agg.most_common(int(os.environ.get("TOP", "60"))):
    if k is None: print(f"{n/tot*100:6.2f}%  <none>"); continue
    print(f"{n/tot*100:6.2f}%  samp {samp[k]:7d}  {k[0]}:{k[1]:5d}  {text(*k)}")
