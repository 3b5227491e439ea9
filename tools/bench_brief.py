"""Lab helper: run bench.py (extra args passed through) and print one short line: value, ms/step, per-kernel ms, parity.
    ORDM_K1ACT_UNDER=1 python tools/bench_brief.py --steps 20 --warmup 3 --lean"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
r = subprocess.run([sys.executable, os.path.join(root, "bench.py")] + sys.argv[1:], capture_output=True, text=True)
if r.returncode:
    print("bench failed:", r.stderr[-1500:])
    sys.exit(1)
d = json.loads(r.stdout.strip().splitlines()[-1])
ms = lambda t: {k: round(v["ms"], 3) for k, v in t.items() if isinstance(v, dict) and "ms" in v}
env = {k: v for k, v in os.environ.items() if k.startswith("ORDM_")}
print(env, "value %.1f M pts/s, %.3f ms/step" % (d["value"] / 1e6, d["ms_per_step"]), ms(d["kernels"]), "isolated", ms(d["kernels"].get("isolated", {})),
      "parity", d["parity"] and d["parity"]["abs_diff"], "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2))
