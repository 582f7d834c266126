#!/bin/bash
# K2 (Wiener-Khinchin path) on one GPU: parity of every plan, the config-5 bench line, then one ncu --set full capture of
# the five passes of one step (second repetition).   gpurun --timeout 1200 -- 'bash tools/gpu_fft_check.sh TAG'
TAG=${1:-fft}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft or method" > $OUT/${TAG}_pytest_fft.log 2>&1
echo "pytest fft rc=$?" | tee -a $OUT/${TAG}_pytest_fft.log; tail -n 8 $OUT/${TAG}_pytest_fft.log
python bench.py --workload fft --no-cpu > $OUT/${TAG}_bench_fft.json 2> $OUT/${TAG}_bench_fft.err; echo "bench fft rc=$?"
cat $OUT/${TAG}_bench_fft.json | head -c 2500; echo
ACFB200_FFT_THREE=0 python bench.py --workload fft --no-cpu > $OUT/${TAG}_bench_fft_pow2.json 2> $OUT/${TAG}_bench_fft_pow2.err
python -c "
import json
for f in ('$OUT/${TAG}_bench_fft.json','$OUT/${TAG}_bench_fft_pow2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['frac'], d['config']['plan'], d['parity'])
"
python tools/one_shot_fft.py 27 2 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"col_pass|row_pass" -s 5 -c 5 -f -o $OUT/${TAG}_prof \
    python tools/one_shot_fft.py 27 2 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 3 $OUT/${TAG}_ncu.log
