#!/bin/bash
# One gpurun call that re-establishes the whole measured state on a fresh B200 box (about 6-8 minutes of box time):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh r2a'
# Everything lands in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
# Order: the parity suite first (a failure there makes the numbers meaningless), then smoke, the bench lines of every
# workload, the reference arm, and -- only after those exited without a profiler -- the ncu launch list of the
# default bench command.
TAG=${1:-check}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -n 3 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_ou.json 2> $OUT/${TAG}_ou.err; echo "bench ou rc=$?"
python bench.py --n 1e8 --steps 5 > $OUT/${TAG}_ou_n1e8.json 2> $OUT/${TAG}_ou_n1e8.err
for w in windows viscosity fft msd; do
    python bench.py --workload $w > $OUT/${TAG}_$w.json 2> $OUT/${TAG}_$w.err; echo "bench $w rc=$?"
done
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err
python tools/cli_compare.py --patched > $OUT/${TAG}_cli.jsonl 2> $OUT/${TAG}_cli.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench_ou.csv \
    python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_ncu.log 2>&1
for f in ou ou_n1e8 windows viscosity fft msd; do
    python - $OUT/${TAG}_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f value %.4g frac %.3f e2e %.4g" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
