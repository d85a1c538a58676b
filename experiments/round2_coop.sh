#!/bin/bash
# Round 2, last measurement: the cooperative scan kernels (32 columns x 8 threads) at config 5,
# where one thread per column is the default.
mkdir -p gpurun_out
for v in default diag_coop b_coop; do
  case $v in
    default) env_="" ;;
    diag_coop) env_="TB200_DIAG_IMPL=coop" ;;
    b_coop) env_="TB200_B_IMPL=coop" ;;
  esac
  env $env_ timeout 60 python bench.py --steps 5 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r02_coop_$v.log 2>&1
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    line = [l for l in open(f"gpurun_out/r02_coop_{v}.log") if l.startswith("{")][-1]
    d = json.loads(line)
    r = d.get("roofline", {})
    ks = {k: round(x["ms_per_launch"], 3) for k, x in r.get("kernels", {}).items()}
    ks.update({k: round(x["ms_per_launch"], 3) for k, x in r.get("other_kernels", {}).items()})
    print(v, "ms/step", round(d["ms_per_step"], 3), ks)
except Exception as exc:
    print(v, "FAILED", exc)
PY
done
