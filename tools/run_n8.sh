#!/bin/bash
# Everything that needs the 8-GPU box, in one gpurun --gpus 8 call (charged 8x: keep it short).
mkdir -p gpurun_out
bash tools/host_facts.sh gpurun_out/r2_host_facts_n8.txt
timeout 300 python tools/pcie_matrix.py --gpus 1,2,4,8 --mb 128 --iters 6 --procs 2,8 > gpurun_out/r2_pcie_matrix_n8.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_group.py tests/test_cuda_memory.py -x -q -m gpu > gpurun_out/r2_pytest_multigpu.txt 2>&1
for n in 2 8; do
  timeout 300 python bench.py --gpus $n --single-process --no-extras --no-classes --steps 10 --e2e-steps 5 \
      > gpurun_out/r2_bench_sp_n$n.json 2> gpurun_out/r2_bench_sp_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --no-extras --no-classes --steps 10 --e2e-steps 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --no-extras --no-classes --steps 10 --e2e-steps 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_pytest_multigpu.txt
tail -c 300 gpurun_out/r2_bench_sp_n8.err gpurun_out/r2_bench_n8.err
