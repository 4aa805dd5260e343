import json, subprocess, sys
p = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
if not lines:
    print("NO JSON. stdout tail:", p.stdout[-2000:], "\nstderr tail:", p.stderr[-3000:])
    sys.exit(1)
d = json.loads(lines[-1])
print("value %.1f Msplats/s  ms/step %.3f  e2e %s  launches %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("gpu_launches")))
r = d.get("roofline") or {}
print("stages ms/step:", {k: round(v, 4) for k, v in (r.get("stage_ms_per_step") or {}).items()})
print("dominant:", r.get("kernel"), "frac", r.get("frac"), "| whole-step frac", (r.get("whole_step") or {}).get("frac"))
print("clocks:", d.get("clocks"), "cpu:", (d.get("cpu_baseline") or {}).get("value"))
