"""Static instruction mix of the innermost loop that contains SHFL.UP in a kernel's SASS.
usage: python scripts/sass_loop.py <mangled-substring>"""
import re, subprocess, sys
from collections import Counter
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", "c2ray3dm_b200/libc2ray_b200.so"], capture_output=True, text=True).stdout
lines = []
on = False
for l in out.splitlines():
    if "Function :" in l:
        on = pat in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(lines)}
shfl = [i for i, (_, ins) in enumerate(lines) if "SHFL.UP" in ins]
best = None
for i, (a, ins) in enumerate(lines):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?`?\(?(0x[0-9a-f]+)", ins)
    if m:
        t = int(m.group(1), 16)
        if t < a and t in addr and any(addr[t] <= s <= i for s in shfl):
            span = i - addr[t]
            if best is None or span < best[2]:
                best = (addr[t], i, span)
if not best:
    print("loop not found; total instr", len(lines)); sys.exit()
body = lines[best[0]:best[1] + 1]
c = Counter()
for _, ins in body:
    parts = ins.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    c[op.split(".")[0] + (".MOV" if "IMAD.MOV" in op else "")] += 1
print("kernel instr %d, loop body %d instr" % (len(lines), len(body)))
fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
print("FP64 pipe:", fp64, " MOV-like:", c.get("IMAD.MOV", 0) + c.get("MOV", 0))
print(", ".join("%s %d" % kv for kv in c.most_common(22)))
