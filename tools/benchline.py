"""stdin: one bench.py JSON line -> compact summary (optionally prefixed by argv[1])."""
import json, sys
d = json.loads(sys.stdin.read())
tag = sys.argv[1] if len(sys.argv) > 1 else ""
print(tag, "depth", d["config"].get("steps_in_flight"), "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"],
      "ms/step %.2f" % d["ms_per_step"], "stages" if "-v" in sys.argv else "", d["roofline"]["stages_ms"] if "-v" in sys.argv else "")
