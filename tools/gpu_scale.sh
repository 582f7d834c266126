#!/bin/bash
# Multi-GPU evidence on one box: the in-library multi-GPU tests, then bench.py under torchrun on all visible GPUs.
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_scale.sh TAG'
TAG=${1:-scale}
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $OUT/${TAG}_topo${NG}.txt 2>&1; nproc >> $OUT/${TAG}_topo${NG}.txt; lscpu | grep -E "NUMA|Model name|Socket" >> $OUT/${TAG}_topo${NG}.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -v > $OUT/${TAG}_pytest_multi_${NG}gpu.log 2>&1
echo "pytest multi rc=$?" | tee -a $OUT/${TAG}_pytest_multi_${NG}gpu.log; tail -n 10 $OUT/${TAG}_pytest_multi_${NG}gpu.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 \
    > $OUT/${TAG}_scale${NG}.out 2> $OUT/${TAG}_scale${NG}.err
echo "bench N=$NG rc=$?"
grep '^{"metric"' $OUT/${TAG}_scale${NG}.out | tail -n 1 > $OUT/${TAG}_scale${NG}.json
grep -E "NVLS|AllReduce|nranks $NG" $OUT/${TAG}_scale${NG}.out $OUT/${TAG}_scale${NG}.err | tail -n 12 > $OUT/${TAG}_scale${NG}_nccl_tail.txt
rm -f $OUT/${TAG}_scale${NG}.out
python - $OUT/${TAG}_scale${NG}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print("ou: ms %.3f value %.4g e2e ms %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
for k, r in d["multi_gpu"].items():
    if "error" in r: print(k, r); continue
    print(k, "ms %.3f" % r["ms_per_step"], "speedup %.3f" % r["speedup_vs_1gpu_same_run"], "e2e ms %.3f" % r["e2e"]["ms_per_step"],
          "e2e speedup %.3f" % r.get("e2e_speedup_vs_1gpu_same_run", float("nan")), "parity", r["parity_max_rel"], "one", r["one_gpu_same_run"])
PY
grep -v "NCCL INFO" $OUT/${TAG}_scale${NG}.err | tail -n 8
