#!/bin/bash
# e2e (host -> host) of bench.py for several HostPipeline chunk sizes, same box
for c in "$@"; do
  python bench.py --steps 20 --no-secondary --no-cpu-baseline --e2e-chunk $c 2>/dev/null > /tmp/e2e_$c.json
  python - $c <<'PY'
import json, sys
c = sys.argv[1]
d = json.load(open(f"/tmp/e2e_{c}.json")); e = d["e2e"]
print("chunk", c, "clips/s", round(e["clips_per_s"], 1), "of ceiling", round(e["frac_of_host_link_ceiling"], 3), "h2d/d2h GB/s",
      round(e["host_link"]["h2d_gbs"], 1), round(e["host_link"]["d2h_gbs"], 1), "| value", round(d["value"]), "frac", round(d["roofline"]["frac"], 3),
      "noise-free", round(d["roofline"]["noise_free_frac"], 3))
PY
done
