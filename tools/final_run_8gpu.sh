# The multi-GPU measurement set of a build on one 8-GPU box; gpurun --gpus 8: bash tools/final_run_8gpu.sh
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 > gpurun_out/r2w_bench_8gpu.json 2> gpurun_out/r2w_bench_8gpu.err
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --bytes 8589934592 --no-e2e --no-cpu > gpurun_out/r2w_bench_cfg5_8gpu.json 2> gpurun_out/r2w_bench_cfg5_8gpu.err
ls -la gpurun_out/r2w*
