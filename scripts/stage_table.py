"""Print ms/step, Msplats/s and the per-stage times of one or more bench.py JSON lines side by side."""
import json, sys
rows = []
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable:", e)
        continue
    rows.append((p, d))
for p, d in rows:
    st = d["roofline"]["stage_ms_per_step"]
    e2e = d.get("e2e") or {}
    print("%s: %.3f ms/step  %.0f Msplats/s  e2e %s  clocks %s" % (p, d["ms_per_step"], d["value"], e2e.get("value"), d.get("clocks")))
    print("   " + "  ".join("%s %.3f" % (k, v) for k, v in st.items()))
