# Final captures of round 2, second session (decoder walk / synthesis rework).  Run on the GPU box from the repo root:
#   bash tools/final_capture_r2b.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/r2b_pytest_gpu.log
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo bench_exit=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_ref.err; echo ref_exit=$?
# decoder: one launch pair (SRLA_B200_DECODE_PIPELINE=0) and the pipelined call, lane counts of the parse kernel
bash tools/decode_sweep.sh > gpurun_out/r2b_decode_sweep.jsonl 2>&1
SRLA_B200_DECODE_PIPELINE=0 python tools/bench_decode.py > gpurun_out/r2b_decode.json 2> gpurun_out/r2b_decode.err; cut -c1-400 gpurun_out/r2b_decode.json
SRLA_B200_DECODE_PIPELINE=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_ -c 8 --csv --log-file gpurun_out/r2b_decode_launches.csv python tools/bench_decode.py > gpurun_out/d_ncu.log 2>&1
SRLA_B200_DECODE_PIPELINE=0 ncu --set full --clock-control none --import-source on -k regex:decode_ -c 2 -o gpurun_out/r2b_decode_full python tools/bench_decode.py > gpurun_out/d_ncu2.log 2>&1
# encoder launch list of the head (unsplit call: one batch on one stream)
SRLA_B200_SPLIT_DEVICE=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --no-cpu-baseline --files 0 --steps 3 --warmup 3 > gpurun_out/b_ncu.log 2>&1
python tools/bench_configs.py > gpurun_out/r2b_other_configs.jsonl 2>&1; tail -c 600 gpurun_out/r2b_other_configs.jsonl
