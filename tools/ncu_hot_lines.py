"""Ranks CUDA source lines of an ncu report by warp-stall samples:  python tools/ncu_hot_lines.py rep.ncu-rep [N]"""
import csv
import subprocess
import sys

txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(txt.splitlines()))
out, hdr, f = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        f = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        n = int(r[si] or 0)
        if n > 0:
            out.append((n, f, int(r[0]), r[1][:100], r))
tot = sum(o[0] for o in out)
print("total samples", tot)
keys = [(i, k) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
for n, f, l, src, r in sorted(out, key=lambda o: -o[0])[:top]:
    st = sorted(((int(r[i] or 0), k[6:]) for i, k in keys), reverse=True)[:3]
    print(f"{n:6d} {100 * n / tot:5.1f}% {f}:{l} {src}  | {st}")
