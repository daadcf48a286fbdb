"""Pretty-print bench.py JSON lines from stdin:  python bench.py | python tools/benchline.py"""
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.0f  ms/step %.3f  e2e %.0f  frac %.4f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
    else: print(l.rstrip())
