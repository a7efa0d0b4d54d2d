# The single-GPU measurement set of a build (bench lines, sweep, ncu launch list, one ncu --set full capture); gpurun: bash tools/final_run.sh
set -x
python bench.py > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench_1gpu.err
python bench.py --impl reference > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err
( python bench.py --mode adaptive ; python bench.py --mode adaptive --chunk 16384 --no-e2e ; python bench.py --mode adaptive --chunk 262144 --no-e2e ; python bench.py --alphabet 4096 ) > gpurun_out/r2u_bench_modes.jsonl 2> gpurun_out/r2u_bench_modes.err
python tools/sweep.py --big --parts 4,8,16 > gpurun_out/r2u_sweep.jsonl 2> gpurun_out/r2u_sweep.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2u_plain.json 2> gpurun_out/r2u_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2u_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"encode_kernel|decode_kernel|hist_global" -s 9 -c 3 -o gpurun_out/r2u_coders -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2u_ncu2.log 2>&1
