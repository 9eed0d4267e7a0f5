#!/bin/bash
# e2e leg of bench.py against the host-pipeline chunk size
for c in 262144 393216 524288 786432 1048576; do
  python bench.py --no-cpu --skip-extra --steps 5 --warmup 3 --e2e-chunk $c 2>/dev/null > /tmp/b_$c.json
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.load(open(f"/tmp/b_{c}.json"))
print("chunk", c, "e2e M rot/s", round(d["e2e"]["value"] / 1e6, 1), "ms", round(d["e2e"]["ms_per_step"], 2))
PY
done
