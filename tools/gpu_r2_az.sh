#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/r2az_fence.log
import os, subprocess, json, sys
for v in ("default", "fence1"):
    env = dict(os.environ)
    if v != "default":
        env["SQOA_B200_LIB"] = os.path.join(os.getcwd(), "gpurun_variants", f"libsqoa_b200_{v}.so")
    out = subprocess.run([sys.executable, "bench.py", "--skip-configs", "--steps", "20", "--warmup", "3"], env=env, capture_output=True, text=True, timeout=120).stdout
    for l in out.splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            print(v, round(d["value"]), {k: round(x["ms"] / 16 * 1000, 1) for k, x in d["legs"].items()}, d.get("parity_spot_check"))
PY
