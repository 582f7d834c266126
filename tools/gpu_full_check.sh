#!/bin/bash
# Single-GPU state check: the whole -m gpu suite, smoke, the default bench line (headline + every config), the FFT line.
#   gpurun --timeout 1500 -- 'bash tools/gpu_full_check.sh TAG'
TAG=${1:-full}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -n 6 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench1.json 2> $OUT/${TAG}_bench1.err; echo "bench rc=$?"; tail -n 5 $OUT/${TAG}_bench1.err
python - $OUT/${TAG}_bench1.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(name, r):
    e = r.get("e2e") or {}
    print("%-10s ms/step %9.3f  value %.4g  frac %.3f  e2e ms %s  parity %s" % (name, r["ms_per_step"], r["value"], (r.get("roofline") or {}).get("frac", float("nan")), e.get("ms_per_step"), r.get("parity")))
show("ou", d)
print("alt", d.get("alt_method"))
for k, r in (d.get("configs") or {}).items():
    if "error" in r: print(k, r)
    else:
        show(k, r)
        if r.get("alt_method"): print("   alt", {kk: vv for kk, vv in r["alt_method"].items() if kk != "parity"})
print("cpu", d.get("cpu_baseline"))
PY
