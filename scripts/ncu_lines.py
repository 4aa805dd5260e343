"""Per-SOURCE-LINE executed warp instructions and stall samples from an .ncu-rep (ncu --import-source on),
top lines only.  usage: ncu_lines.py report.ncu-rep [min_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, hdr, acc, idx = None, None, [], {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        iE, iN = hdr.index("Instructions Executed"), hdr.index("# Samples")
        iT = hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > iE and r[iE].isdigit() and r[0].isdigit():
        key = (cur_file, int(r[0]))
        if key not in idx:
            idx[key] = len(acc)
            acc.append([cur_file, int(r[0]), r[1].strip(), 0, 0, 0])
        a = acc[idx[key]]
        a[3] += int(r[iE]); a[4] += int(r[iN]) if r[iN].isdigit() else 0; a[5] += int(r[iT]) if r[iT].isdigit() else 0
totE = sum(a[3] for a in acc) or 1
totN = sum(a[4] for a in acc) or 1
print("total warp instructions %d, stall samples %d" % (totE, totN))
for f, ln, src, e, n, t in acc:
    if 100.0 * e / totE >= minpct or 100.0 * n / totN >= minpct:
        print("%-16s %4d  exec %5.2f%%  samp %5.2f%%  thr/inst %4.1f  %s" % (f, ln, 100.0 * e / totE, 100.0 * n / totN, t / max(e, 1), src[:90]))
