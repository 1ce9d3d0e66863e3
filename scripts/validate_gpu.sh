#!/usr/bin/env bash
# Round-end validation on a B200 box (everything bounded by `timeout`; writes into gpurun_out/).
#   gpurun --timeout 1500 -- 'bash scripts/validate_gpu.sh 1'          # one GPU: tests, bench, launch list, ncu capture
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/validate_gpu.sh 2'  # two GPUs: multi-GPU tests + sharded bench
# Never sets CVD_ATTN_CLUSTER on a multi-GPU run: multicast TMA from peer memory wedges the GPU.
set -u
N=${1:-1}
mkdir -p gpurun_out
unset CVD_ATTN_CLUSTER
if [ "$N" = "1" ]; then
  timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/v_tests.log; tail -3 gpurun_out/v_tests.log
  timeout -k 5 700 python bench.py --steps 10 --warmup 3 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"
  CVD_ATTN_CLUSTER=1 timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-denoise --no-extras > gpurun_out/v_bench_cluster.json 2>/dev/null
  timeout -k 5 200 python bench.py --steps 2 --warmup 1 --no-denoise --no-extras > gpurun_out/v_plain.log 2>&1 &&
    timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-denoise --no-extras > gpurun_out/v_ncu_list.log 2>&1
else
  timeout -k 5 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/v_mg_tests.log; tail -3 gpurun_out/v_mg_tests.log
  for mode in "" "--no-graph"; do
    timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus "$N" --steps 10 --warmup 3 --no-denoise $mode > "gpurun_out/v_bench_${N}gpu${mode}.json" 2> "gpurun_out/v_bench_${N}gpu${mode}.err"
    echo "bench N=$N $mode rc=$?"
  done
fi
