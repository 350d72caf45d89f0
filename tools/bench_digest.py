"""One line per bench JSON: step time, value, e2e, probe times (us) and the two roofline fractions."""
import json
import sys

for f in sys.argv[1:]:
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    km = {k: round(v * 1e3, 1) for k, v in j.get("kernel_ms", {}).items() if v is not None}
    print("%s: %.3f ms/step  value %.0f  e2e %.0f  %s  gather %.3f  scatter %.3f" % (
        f.split("/")[-1], j["ms_per_step"], j["value"], j["e2e"]["value"], km, (j.get("roofline_gather") or j["roofline"])["frac"] or 0,
        j["roofline_scatter"]["frac"] or 0))
