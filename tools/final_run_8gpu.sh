# The multi-GPU measurement set of a build on one 8-GPU box; gpurun --gpus 8: bash tools/final_run_8gpu.sh
for n in 4 8; do
  NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n > gpurun_out/r2v_bench_${n}gpu.json 2> gpurun_out/r2v_bench_${n}gpu.err
done
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --bytes 8589934592 --no-e2e --no-cpu > gpurun_out/r2v_bench_cfg5_8gpu.json 2> gpurun_out/r2v_bench_cfg5_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --impl reference --gpus 8 > gpurun_out/r2v_bench_ref_8gpu.json 2> gpurun_out/r2v_bench_ref_8gpu.err
ls -la gpurun_out/r2v*
