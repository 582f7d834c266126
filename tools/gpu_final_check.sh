#!/bin/bash
# Final single-GPU state of the round: the -m gpu suite (direct method and auto method), smoke, the default bench line, the
# reference arm, and the ncu launch list of the default bench command.
#   gpurun --timeout 1800 -- 'bash tools/gpu_final_check.sh TAG'
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log; tail -n 4 $OUT/${TAG}_pytest.log
ACFB200_TEST_METHOD=auto timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_auto.log 2>&1; echo "pytest auto rc=$?" | tee -a $OUT/${TAG}_pytest_auto.log; tail -n 3 $OUT/${TAG}_pytest_auto.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench1.json 2> $OUT/${TAG}_bench1.err; echo "bench rc=$?"; tail -n 3 $OUT/${TAG}_bench1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err; echo "reference rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench_ou.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
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
