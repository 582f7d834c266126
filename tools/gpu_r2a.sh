#!/bin/bash
# Round-2 first hardware pass on >= 2 GPUs: the in-library multi-GPU tests (kept log), the multi-GPU bench line with
# its windows / viscosity sub-records, then the single-GPU suite and the single-GPU bench line with every config.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r2a.sh r2a'
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/${TAG}_smi.txt 2>&1
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
nproc >> $OUT/${TAG}_topo.txt; numactl -H >> $OUT/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -v > $OUT/${TAG}_pytest_multi_${NG}gpu.log 2>&1
echo "pytest multi rc=$?" | tee -a $OUT/${TAG}_pytest_multi_${NG}gpu.log; tail -n 12 $OUT/${TAG}_pytest_multi_${NG}gpu.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 \
    > $OUT/${TAG}_scale${NG}.json 2> $OUT/${TAG}_scale${NG}.err
echo "bench N=$NG rc=$?"; tail -c 3000 $OUT/${TAG}_scale${NG}.json; grep -v "NCCL INFO" $OUT/${TAG}_scale${NG}.err | tail -n 20
grep -E "AllReduce|NVLS|Connected all|comm 0x.* rank 0 nranks" $OUT/${TAG}_scale${NG}.err | tail -n 12 > $OUT/${TAG}_scale${NG}_nccl_tail.txt
timeout 600 python bench.py > $OUT/${TAG}_bench1.json 2> $OUT/${TAG}_bench1.err
echo "bench N=1 rc=$?"; tail -c 6000 $OUT/${TAG}_bench1.json; tail -n 20 $OUT/${TAG}_bench1.err
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -n 5 $OUT/${TAG}_pytest.log
