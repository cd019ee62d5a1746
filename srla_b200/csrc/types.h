/*
 * types.h -- descriptors shared by the host runtime and the CUDA kernels of libsrla_b200.
 * Vocabulary: a *stream* is one audio file, a *job* is one block (or candidate segment of the
 * variable-block search) of a stream, a *candidate* is one channel signal a block may be coded
 * from (for stereo: M, S, L, R), exactly the four calls the reference makes per block
 * (srla_encoder.c:1248-1273).
 */
#ifndef SRLA_B200_TYPES_H
#define SRLA_B200_TYPES_H

#include <stdint.h>

namespace srla {

constexpr int kThreads        = 256;   /* threads per CTA in every kernel                      */
constexpr int kMaxOrder       = 255;   /* SRLA_MAX_COEFFICIENT_ORDER                           */
constexpr int kMaxChannels    = 8;
constexpr int kMaxCand        = kMaxChannels + 2;
constexpr int kMaxBlock       = 16384; /* capacity of the shared-memory resident pipeline       */
constexpr int kLog2MaxParts   = 10;    /* srla_coder.c:18                                       */
constexpr int kMaxParts       = 1 << kLog2MaxParts;
constexpr int kLtpMinPeriod   = 8;     /* srla_internal.h:31-35                                 */
constexpr int kLtpMaxPeriod   = 8 + 256 - 2;
constexpr int kLtpLags        = kLtpMaxPeriod + 3;  /* lags the pitch search may touch (0..264)   */

enum BlockType { kBlockCompress = 0, kBlockSilent = 1, kBlockRaw = 2 };
enum CodeType  { kCodeRice = 0, kCodeRecursiveRice = 1, kCodeAllZero = 2 };

/* one audio file resident in HBM, planar: channel c at pcm + c * stride (samples) */
struct StreamDev {
    const void *pcm;
    unsigned long long stride;
    uint32_t num_samples;
    uint32_t sample_bytes;     /* 2: int16_t, 4: int32_t                                  */
    uint32_t lshift;           /* common trailing-zero shift (written by lshift_finish_kernel or host) */
    uint32_t or_mask;          /* scratch of the OR-reduction                              */
};

/* one block to analyse / emit */
struct Job {
    uint32_t stream;           /* index into StreamDev[]                                   */
    uint32_t offset;           /* first sample (per channel) inside the stream             */
    uint32_t nsmpl;            /* samples per channel                                      */
    uint32_t flags;            /* kJobFirstOfStream: the 30-byte stream header precedes it */
    double   welch_div;        /* 4 * pow(n-1, -2), host libm (lpc.c:259)                  */
    double   welch_gain;       /* window energy compensation (lpc.c:275-290)               */
    double   ac_scale;         /* 2.0 / n (lpc.c:336)                                      */
};
constexpr uint32_t kJobFirstOfStream = 1u;

/* result of analysing one candidate channel of one job */
struct CandOut {
    int32_t  pre_coef, pre_prev;
    uint32_t order, rshift, use_sum, coef_bits;
    uint32_t ltp_period;
    int32_t  ltp_coef[3];
    uint32_t code_type, porder, residual_bits, total_bits;
    uint32_t nonzero;          /* raw input of this channel has a non-zero sample          */
    uint32_t status;           /* 0 ok, 1: reference would fail the encode (singular LTP system) */
    int16_t  coef[256];        /* FIR order                                                */
    uint8_t  kparam[kMaxParts];/* coding parameter per partition at `porder`               */
};

/* block-level decision */
struct JobOut {
    uint32_t type;             /* BlockType                                                */
    uint32_t method;           /* 0 LR, 1 MS, 2 LS, 3 SR                                   */
    uint32_t bytes;            /* encoded block size incl. the 11-byte block header        */
    uint32_t estimate_bytes;   /* what SRLAEncoder_ComputeBlockSize reports (channels 0/1 only, srla_encoder.c:1276-1301) */
    uint32_t cand_of_channel[kMaxChannels];
    uint32_t status;
    uint32_t pad;
    unsigned long long out_offset;   /* byte offset of the block inside the output buffer   */
};

/* parameters shared by all jobs of a launch */
struct LaunchParams {
    const StreamDev *streams;
    const Job       *jobs;
    CandOut         *cand;           /* [job][cand]                                        */
    JobOut          *jobout;         /* [job]                                              */
    int32_t         *residual;       /* [job][cand][res_stride]                            */
    uint32_t num_jobs;
    uint32_t nch, ncand, bps;
    uint32_t max_order;              /* preset's maximum LPC order                         */
    uint32_t ltp_order;
    uint32_t res_stride;             /* samples                                            */
    uint32_t nmax;                   /* longest job of this launch                         */
    uint32_t fft_max;                /* next power of two >= nmax                          */
    uint32_t sampling_rate, max_block, preset;   /* stream header fields                   */
    /* tables in device memory */
    const double2 *tw_complex;       /* per stage size: {w1,w2,w3}[ns/4]                   */
    const uint32_t *tw_complex_off;  /* [log2(ns)] -> offset in double2 units              */
    const double2 *tw_real;          /* per real size N: {wr,wi}[N/4] forward; inverse conjugates wi */
    const uint32_t *tw_real_off;     /* [log2(N)]                                          */
    const double  *rice_threshold;   /* [32] smallest mean with k >= j (plain Rice)        */
    /* output */
    uint8_t  *out;
    unsigned long long out_capacity;
    unsigned long long *running;     /* [0] bytes emitted so far, [1] overflow flag        */
    unsigned long long *stream_begin;/* [num_streams+1] byte offset where each stream starts */
    uint32_t *stats;                 /* order[256], method[4], type[3]                      */
    uint32_t emit_stream_header;
};

} // namespace srla
#endif
