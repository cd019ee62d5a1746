#!/bin/bash
# Decode throughput on the config-2 stream for several lane counts of decode_parse_kernel, one-launch and pipelined calls.
# Usage (on the GPU box): bash tools/decode_sweep.sh > gpurun_out/decode_sweep.jsonl
for pipe in 0 1; do
  for lanes in 0 4 8 16 32; do
    echo "# SRLA_B200_DECODE_PIPELINE=$pipe SRLA_B200_DECODE_LANES=$lanes"
    SRLA_B200_DECODE_PIPELINE=$pipe SRLA_B200_DECODE_LANES=$lanes python tools/bench_decode.py
  done
done
