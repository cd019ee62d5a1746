"""srla_b200 -- B200-native (sm_100a CUDA) SRLA lossless-audio encode path.

csrc/            CUDA kernels + the extern "C" shim (libsrla_b200.so, built by __graft_entry__.build())
encoder.py       ctypes mirror of the reference's SRLAEncoder_* interface + the batch extension
synth.py         deterministic synthetic PCM (SURVEY.md 8d recipe)
workload.py      BASELINE.json benchmark workloads
sharding.py      multi-GPU partition of a batch of streams
"""
