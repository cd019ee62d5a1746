mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo bench_exit=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2> gpurun_out/r2_ref.err; echo ref_exit=$?
# launch list and full capture of the UNSPLIT call (one batch on one stream: the same kernels, not overlapped)
SRLA_B200_SPLIT_DEVICE=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --no-cpu-baseline --files 0 --steps 3 --warmup 3 > gpurun_out/b_ncu.log 2>&1
SRLA_B200_SPLIT_DEVICE=0 ncu --set full --clock-control none --import-source on -s 18 -c 9 -o gpurun_out/r2_final_full python bench.py --no-cpu-baseline --files 0 --steps 2 --warmup 2 > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out/r2_final_full.ncu-rep
python tools/bench_decode.py > gpurun_out/r2_decode.json 2> gpurun_out/r2_decode.err; tail -2 gpurun_out/r2_decode.json | cut -c1-600
