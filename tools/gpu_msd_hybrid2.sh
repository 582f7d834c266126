#!/bin/bash
# Second short GPU call for the hybrid MSD: the front-end's --method extension, the timing line with the end-to-end
# leg, and the same at --max 1e6.
OUT=gpurun_out
mkdir -p $OUT
timeout 60 python -m pytest tests/test_cli.py -m gpu -q -x -k "method_auto" > $OUT/r2_msd_hybrid_cli_pytest.log 2>&1
echo "pytest cli rc=$?" | tee -a $OUT/r2_msd_hybrid_cli_pytest.log; tail -n 25 $OUT/r2_msd_hybrid_cli_pytest.log
timeout 40 python tools/msd_hybrid_bench.py > $OUT/r2_msd_hybrid_bench_1e5.json 2> $OUT/r2_msd_hybrid_bench.err
echo "bench rc=$?"; cat $OUT/r2_msd_hybrid_bench_1e5.json; tail -n 5 $OUT/r2_msd_hybrid_bench.err
timeout 60 python tools/msd_hybrid_bench.py --max 1e6 --reps 1 > $OUT/r2_msd_hybrid_bench_1e6.json 2> $OUT/r2_msd_hybrid_bench.err
echo "bench rc=$?"; cat $OUT/r2_msd_hybrid_bench_1e6.json; tail -n 5 $OUT/r2_msd_hybrid_bench.err
