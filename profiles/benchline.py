import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print("%.4g steps/s  %.3f ms  frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
