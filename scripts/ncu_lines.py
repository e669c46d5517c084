"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) // 2:
        continue
    if r[0] != "":  # a source line row
        lines.append(r)
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


iS = hdr.index("# Samples")
iI = hdr.index("Instructions Executed")
iT = hdr.index("Thread Instructions Executed")
tot_i = sum(num(r[iI]) for r in lines)
tot_s = sum(num(r[iS]) for r in lines)
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("%6s %7s %7s  %s" % ("line", "inst%", "smpl%", "source"))
key = (lambda r: -num(r[iS])) if len(sys.argv) > 3 and sys.argv[3] == 'samples' else (lambda r: -num(r[iI]))
for r in sorted(lines, key=key)[:top]:
    print("%6s %6.2f%% %6.2f%%  %s" % (r[0], 100.0 * num(r[iI]) / tot_i, 100.0 * num(r[iS]) / max(1, tot_s), r[1][:110]))
