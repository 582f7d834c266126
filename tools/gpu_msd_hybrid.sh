#!/bin/bash
# One short GPU call for the hybrid Wiener-Khinchin MSD: its tests (library and front-end) first, then the timing line, then
# the rest of the MSD file.
#   gpurun --timeout 120 -- 'bash tools/gpu_msd_hybrid.sh'
OUT=gpurun_out
mkdir -p $OUT
timeout 100 python -m pytest tests/test_msd.py tests/test_cli.py -m gpu -q -x -k "hybrid" > $OUT/r2_msd_hybrid_pytest.log 2>&1
echo "pytest hybrid rc=$?" | tee -a $OUT/r2_msd_hybrid_pytest.log; tail -n 25 $OUT/r2_msd_hybrid_pytest.log
timeout 60 python tools/msd_hybrid_bench.py > $OUT/r2_msd_hybrid_bench.json 2> $OUT/r2_msd_hybrid_bench.err
echo "bench rc=$?"; cat $OUT/r2_msd_hybrid_bench.json; tail -n 5 $OUT/r2_msd_hybrid_bench.err
timeout 100 python -m pytest tests/test_msd.py -m gpu -q -x -k "not hybrid" > $OUT/r2_msd_rest_pytest.log 2>&1
echo "pytest rest rc=$?" | tee -a $OUT/r2_msd_rest_pytest.log; tail -n 5 $OUT/r2_msd_rest_pytest.log
