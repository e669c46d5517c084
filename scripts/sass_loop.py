"""Static instruction mix of the ray tracer's row loops from cuobjdump -sass (run here, no GPU needed).
usage: python scripts/sass_loop.py [mangled-substring] [--dump FILE]
A row loop = an innermost backward-branch loop that contains a RED.E.ADD.F64 (the rate accumulation) -- one per
inlined trace_shell instantiation (planes in shared/global memory x clipped/unclipped x r==1)."""
import re, subprocess, sys
from collections import Counter
pat = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "raytrace_kernelILi1ELi1ELb0ELb0"
dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
out = subprocess.run(["cuobjdump", "-sass", "c2ray3dm_b200/libc2ray_b200.so"], capture_output=True, text=True).stdout
lines, on, fname = [], False, ""
for l in out.splitlines():
    if "Function :" in l:
        on = pat in l
        if on:
            fname = l.split(":")[1].strip()
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(lines)}
loops = []
for i, (a, ins) in enumerate(lines):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?`?\(?(0x[0-9a-f]+)", ins)
    if m:
        t = int(m.group(1), 16)
        if t < a and t in addr:
            loops.append((addr[t], i))
inner = [lp for lp in loops if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in loops)]
def opname(ins):
    parts = ins.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    base = op.split(".")[0]
    if "IMAD.MOV" in op:
        base = "IMAD.MOV"
    if op.startswith("RED"):
        base = "RED"
    if op.startswith("LDS") or op.startswith("STS") or op.startswith("LDG") or op.startswith("STG"):
        base = op.split(".")[0] + ("." + op.split(".")[-1] if op.split(".")[-1] in ("64", "128") else "")
    return base
print("kernel %s: %d SASS instructions, %d innermost loops" % (fname[-60:], len(lines), len(inner)))
rows = []
for lo, hi in inner:
    body = lines[lo:hi + 1]
    if not any(ins.split()[0].startswith("RED") or (len(ins.split()) > 1 and ins.split()[1].startswith("RED")) for _, ins in body):
        continue
    c = Counter(opname(ins) for _, ins in body)
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    kind = "global planes" if any(k.startswith("LDG.128") or k == "LDG.128" for k in c) else "shared planes"
    rows.append((lo, hi, body, c, fp64, kind))
for lo, hi, body, c, fp64, kind in rows:
    print("-" * 100)
    print("row loop @0x%04x-0x%04x (%s): %d instructions, FP64 pipe %d (DFMA %d DMUL %d DADD %d DSETP %d), MUFU %d, RED %d" % (
        lines[lo][0], lines[hi][0], kind, len(body), fp64, c["DFMA"], c["DMUL"], c["DADD"], c["DSETP"], c["MUFU"], c["RED"]))
    print("  " + ", ".join("%s %d" % kv for kv in c.most_common(30)))
if dump and rows:
    lo, hi, body, c, fp64, kind = max(rows, key=lambda r: len(r[2]))
    with open(dump, "w") as f:
        f.write("# cuobjdump -sass of %s\n# the largest row loop (%s), %d instructions; mix above from scripts/sass_loop.py\n" % (fname, kind, len(body)))
        for a, ins in body:
            f.write("/*%04x*/ %s ;\n" % (a, ins))
