"""Top CUDA source lines of an .ncu-rep by warp-stall samples (needs --import-source on and -lineinfo).
Usage: python tools/ncu_hot_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
lines = []; H = None; fname = ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; continue
    if H is None or r[0] == "Function Name": continue
    if r[0] != "":                     # a CUDA source line: aggregate of its SASS
        sa = H.index("# Samples"); ie = H.index("Instructions Executed")
        try: n = float(r[sa])
        except ValueError: continue
        stalls = []
        for c, i in [(c, i) for i, c in enumerate(H) if c.startswith("stall_") and "Not Issued" not in c]:
            try: stalls.append((float(r[i]), c[6:]))
            except (ValueError, IndexError): pass
        stalls.sort(reverse=True)
        # the CSV splits the source text on commas: re-join the cells before "Address"
        src = ",".join(r[1:len(r) - (len(H) - 2)]) if len(r) > len(H) else r[1]
        lines.append((n, fname, r[0], src.strip()[:100], r[ie] if ie < len(r) else "", stalls[:3]))
tot = sum(l[0] for l in lines)
lines.sort(reverse=True)
print("total samples", tot)
for n, f, ln, src, ie, st in lines[:top]:
    print(f"{100 * n / max(tot, 1):5.1f}% {f}:{ln:>4} " + " ".join(f"{c}:{v:.0f}" for v, c in st) + f" | {src}")
