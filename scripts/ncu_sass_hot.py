"""Per-SASS-instruction executed counts / stall samples from an .ncu-rep (source page), grouped in address ranges."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iA, iS, iN, iE, iT = (hdr.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
tot_e = sum(int(r[iE]) for r in rows[2:] if len(r) > iE)
tot_s = sum(int(r[iN]) for r in rows[2:] if len(r) > iE)
print("total inst executed %d, samples %d" % (tot_e, tot_s))
base = int(rows[2][iA], 16)
for r in rows[2:]:
    if len(r) <= iE: continue
    e, s = int(r[iE]), int(r[iN])
    if top and e < tot_e / top and s < tot_s / top: continue
    print("%05x %-70s exec %6.2f%% samp %6.2f%% thr %s" % (int(r[iA], 16) - base, r[iS].strip()[:70], 100.0 * e / tot_e, 100.0 * s / max(tot_s, 1), r[iT]))
