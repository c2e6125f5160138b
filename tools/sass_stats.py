"""Static look at the SASS of libacmmp_b200.so: for every kernel, registers / stack, and for every
backward-branch loop that contains TEX instructions the instruction count per TEX.  Development aid
(the B200 is TEX-issue bound at 1 warp-TEX per 8 clk per SM = 32 issue slots per TEX at 4 IPC)."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "acmmp-spherical_b200" / "lib" / "libacmmp_b200.so")
want = sys.argv[2] if len(sys.argv) > 2 else "k_pass"

out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if want not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print(f"== {name}: {len(ins)} instructions")
    loops = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`?\(?\.?(0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t in addr_index and addr_index[t] <= i:
                loops.append((addr_index[t], i))
    for a, b in loops:
        body = [s for _, s in ins[a:b + 1]]
        ntex = sum(1 for s in body if re.search(r"\bTEX\b|\bTEX\.", s))
        if ntex == 0:
            continue
        # skip loops that merely wrap inner loops with TEX (report innermost only)
        inner = [(c, d) for c, d in loops if (c, d) != (a, b) and c >= a and d <= b and
                 any(re.search(r"\bTEX", s) for _, s in ins[c:d + 1])]
        tag = "outer" if inner else "INNER"
        ops = {}
        for s in body:
            op = s.split()[0]
            if op.startswith("@"):
                op = s.split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14])
        print(f"  loop [{a}..{b}] {tag} n={b - a + 1} TEX={ntex} instr/TEX={(b - a + 1) / ntex:.1f}  {top}")
