#!/bin/bash
# Shorter form of gpu_final_check.sh for a change that touches only the Wiener-Khinchin kernels: the -m gpu suite (direct
# and auto method), smoke, the default bench line, and one ncu --set full capture of the five passes of a config-5 step.
#   gpurun --timeout 900 -- 'bash tools/gpu_final_check2.sh TAG'
TAG=${1:-final2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log; tail -n 4 $OUT/${TAG}_pytest.log
ACFB200_TEST_METHOD=auto timeout 600 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_auto.log 2>&1; echo "pytest auto rc=$?" | tee -a $OUT/${TAG}_pytest_auto.log; tail -n 3 $OUT/${TAG}_pytest_auto.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench1.json 2> $OUT/${TAG}_bench1.err; echo "bench rc=$?"; tail -n 3 $OUT/${TAG}_bench1.err
python - $OUT/${TAG}_bench1.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{"metric"')][-1])
def show(name, r):
    e = r.get("e2e") or {}
    p = r.get("parity") or {}
    print("%-10s ms/step %9.3f  value %.4g  frac %.3f nominal %s  e2e ms %.3f  parity %s" % (name, r["ms_per_step"], r["value"], (r.get("roofline") or {}).get("frac", float("nan")), (r.get("roofline") or {}).get("frac_of_nominal"), e.get("ms_per_step", float("nan")), p.get("max_rel_to_c0", p.get("max_rel"))))
show("ou", d)
print("clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
print("alt", {k: v for k, v in (d.get("alt_method") or {}).items() if k != "parity"})
for k, r in (d.get("configs") or {}).items():
    if "error" in r: print(k, r)
    else:
        show(k, r)
        if r.get("alt_method"): print("   alt", {kk: vv for kk, vv in r["alt_method"].items() if kk != "parity"})
print("cpu", d.get("cpu_baseline"))
PY
python tools/one_shot_fft.py 27 2 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"col_pass|row_pass" -s 5 -c 5 -f -o $OUT/${TAG}_prof \
    python tools/one_shot_fft.py 27 2 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 2 $OUT/${TAG}_ncu.log
