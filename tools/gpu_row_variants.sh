#!/bin/bash
# Times the config-5 Wiener-Khinchin step with several builds of the library (acfcalculator_b200/variants/*.so, same ABI,
# selected through ACFB200_LIB); each bench line carries its own oracle parity.
#   gpurun --timeout 600 -- 'bash tools/gpu_row_variants.sh TAG'
TAG=${1:-rowvar}
OUT=gpurun_out
mkdir -p $OUT
for lib in acfcalculator_b200/variants/*.so; do
    name=$(basename $lib .so)
    ACFB200_LIB=$PWD/$lib timeout 300 python bench.py --workload fft --no-cpu --no-extras --steps 30 --warmup 5 \
        > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
    echo "$name rc=$?"
    python - $OUT/${TAG}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{"metric"')][-1])
    print("   ms %.4f frac %.4f parity %s" % (d["ms_per_step"], d["roofline"]["frac"], d.get("parity")))
except Exception as e:
    print("   no line:", e)
PY
done
for lib in acfcalculator_b200/variants/*.so; do
    name=$(basename $lib .so)
    ACFB200_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"col_pass|row_pass" --csv \
        --log-file $OUT/${TAG}_${name}_launches.csv python tools/one_shot_fft.py 27 2 > $OUT/${TAG}_${name}_ncu.log 2>&1
    echo "$name per-pass us (second repetition):"
    python - $OUT/${TAG}_${name}_launches.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
print("   ", [(r[4][:14], r[-1]) for r in rows[-5:]])
PY
done
