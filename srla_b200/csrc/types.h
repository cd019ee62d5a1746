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
#include <vector_types.h>   /* double2 (CUDA toolkit header, host-safe) */

namespace srla {

constexpr int kThreads        = 256;   /* threads per CTA in the analyse / emit kernels        */
constexpr int kWarps          = kThreads / 32;
constexpr int kMaxOrder       = 255;   /* SRLA_MAX_COEFFICIENT_ORDER                           */
constexpr int kMaxChannels    = 8;
constexpr int kMaxCand        = kMaxChannels + 2;
constexpr int kMaxBlock       = 65535; /* SRLA block header: u16 sample count (srla_encoder.c:1593)                      */
constexpr int kMaxSharedBlock = 16384; /* capacity of the shared-memory resident pipeline; longer blocks work in HBM   */
constexpr int kLog2MaxParts   = 10;    /* srla_coder.c:18                                       */
constexpr int kMaxParts       = 1 << kLog2MaxParts;
constexpr int kLtpMinPeriod   = 8;     /* srla_internal.h:31-35                                 */
constexpr int kLtpMaxPeriod   = 8 + 256 - 2;
constexpr int kLtpLags        = kLtpMaxPeriod + 4;  /* lags the pitch search may touch (0..265)  */

enum BlockType { kBlockCompress = 0, kBlockSilent = 1, kBlockRaw = 2 };
enum CodeType  { kCodeRice = 0, kCodeRecursiveRice = 1, kCodeAllZero = 2 };

/* one audio file resident in HBM, planar: channel c at pcm + c * stride (samples) */
struct StreamDev {
    const void *pcm;
    unsigned long long stride;
    uint32_t num_samples;
    uint32_t sample_bytes;     /* 2: int16_t, 4: int32_t                                  */
    uint32_t lshift;           /* common trailing-zero shift (srla_utility.c:177-203)      */
    uint32_t or_mask;          /* scratch of the OR-reduction                              */
    const void *raw;           /* WAV ingest: the stream's frames as they lie in a WAV data chunk (interleaved,
                                  little endian), or NULL; deinterleave_jobs_kernel turns them into `pcm`   */
    uint32_t container_bytes;  /* bytes per sample in `raw`: 1 (unsigned, offset 128), 2, 3 (signed)        */
    uint32_t pad_;
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
    /* code bits (parameter fields included) of the samples [t * n / 256, (t + 1) * n / 256) at the chosen partition order,
     * t = 0..255, saturated at 0xffff: the residual stage knows them, so emit_kernel's sizing pass only reads them.
     * Valid when tb_valid (blocks of 1024 * {1,2,3,4,8} samples, i.e. the register-resident Rice search ran). */
    uint16_t thread_bits[kThreads];
    uint32_t tb_valid, pad_tb[3];
};
static_assert(sizeof(CandOut) % 16 == 0, "candidate records are fetched in 16-byte pieces");

/* optional diagnostics of one candidate (stage-level parity tests) */
struct CandDiag {
    double autocorr[kMaxOrder + 1];   /* after the ridge scaling of lag 0 */
    double error_vars[kMaxOrder + 1]; /* window-compensated               */
    double lpc_double[kMaxOrder + 1]; /* un-quantised coefficients of the chosen order */
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

/* front16_kernel (one CTA per job, 16-bit PCM without LTP, transforms of at most 4096 points): the FFT buffer, the int16
 * rows of ALL channels of the job (16 bytes of padding in front of and behind each row; fetched by bulk asynchronous
 * copies), reduction scratch, the candidates' pre-emphasis coefficients, the copies' mbarrier */
struct Front16Layout { uint32_t region_off, rows_off, row_bytes, red_off, coef_off, bar_off, total; };
/* parameters shared by all jobs of a launch */
struct LaunchParams {
    StreamDev       *streams;
    const Job       *jobs;
    CandOut         *cand;           /* [job][cand]                                        */
    CandDiag        *diag;           /* [job][cand] or NULL                                */
    JobOut          *jobout;         /* [job]                                              */
    int32_t         *residual;       /* [job][cand][res_stride] or NULL (size-only pass)   */
    double          *lags;           /* [job][cand][lag_stride]: autocorrelation lags 0..P  */
    uint32_t lag_stride;
    double          *lpc_state;      /* [candidate group of 32][2][P+2][32]: reflection coefficients, error variances */
    double          *svr_coef;       /* SVR refinement only: [job][cand][P] un-quantised coefficients of the chosen order */
    double          *svr_matrix;     /* SVR refinement only: [CTA of svr_kernel][P][P] covariance / Cholesky factor      */
    uint32_t svr_iterations;         /* num_svr_filter_learning_iteration (0: off)                                       */
    uint32_t replay_tails;           /* big-block path: the stale-scratch replay applies (fixed blocks, even block size)  */
    uint32_t group_first;            /* big-block path: index of jobs[0] inside jobs_all                                  */
    uint32_t serial_streams;         /* big-block path: ODD block size -- the calls of a stream form one chain (front_big_kernel) */
    const Job *jobs_all;             /* big-block path: the call's whole job list (a tail's predecessor may lie in an earlier launch) */
    unsigned char *big_scratch;      /* big-block path (blocks beyond the shared-memory capacity): per-CTA scratch in HBM  */
    unsigned long long big_stride;   /* bytes per CTA there                                                               */
    uint32_t num_jobs, num_streams;
    uint32_t nch, ncand, bps;
    uint32_t max_order;              /* preset's maximum LPC order                         */
    uint32_t ltp_order;
    uint32_t res_stride;             /* samples                                            */
    uint32_t nmax;                   /* longest job of this launch                         */
    uint32_t fft_max;                /* next power of two >= nmax                          */
    uint32_t sampling_rate, max_block, preset;   /* stream header fields                   */
    uint32_t fixed_lshift;           /* used when use_fixed_lshift (single-block API)       */
    uint32_t use_fixed_lshift;
    uint32_t emit_stream_header;
    uint32_t emit_smem_bytes;        /* staging capacity of the emit kernel                 */
    double   unit;                   /* 2^-(bps-1)                                         */
    /* tables in device memory */
    const double2  *tw_complex;      /* per stage size: {w1,w2,w3}[ns/4]                   */
    const double2  *tw_real;         /* per real size N: {wr,wi}[N/4] forward; inverse conjugates wi */
    uint32_t tw_complex_off[20];     /* [log2(ns)] -> offset in double2 units              */
    uint32_t tw_real_off[20];        /* [log2(N)]                                          */
    Front16Layout f16;               /* front16_kernel's shared-memory layout (computed by the host: the offsets cost no registers) */
    const double   *rice_threshold;  /* [32] smallest mean with k >= j (plain Rice)        */
    const uint32_t *huff_code;       /* [2][256] plain, summed                             */
    const uint8_t  *huff_len;        /* [2][256]                                           */
    /* output */
    uint8_t  *out;
    unsigned long long out_capacity;
    unsigned long long *running;     /* [0] bytes emitted so far, [1] overflow flag        */
    unsigned long long *stream_begin;/* [num_streams+1] byte offset where each stream starts */
    uint32_t *stats;                 /* order[256], method[4], type[3]                      */
};

/* shared-memory layouts (identical on host and device) */
#if defined(__CUDACC__)
#define SRLA_HD __host__ __device__
#else
#define SRLA_HD
#endif

SRLA_HD inline uint32_t round_up_u32(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

/* front_kernel: FFT buffer, the pre-emphasised signal, LTP lags */
struct FrontLayout {
    uint32_t region_off, region_bytes; /* FFT buffer (doubles) / int32 scratch                 */
    uint32_t sig_off;                  /* int32: 4 pad + nmax rounded up to 4 + 12             */
    uint32_t lags_off, nlags;          /* doubles (LTP pitch search only)                      */
    uint32_t total;
};
SRLA_HD inline FrontLayout make_front_layout(uint32_t nmax, uint32_t fft_max, uint32_t ltp)
{
    FrontLayout L;
    const uint32_t n4 = round_up_u32(nmax, 4);
    uint32_t off = 0;
    L.region_off = off;
    L.region_bytes = round_up_u32((8u * fft_max > 4u * n4) ? 8u * fft_max : 4u * n4, 16);
    off += L.region_bytes;
    L.sig_off = off; off += 4u * (n4 + 16u);
    L.nlags = ltp ? round_up_u32((uint32_t)kLtpLags, 2) : 0u;
    L.lags_off = off; off += 8u * L.nlags;
    L.total = off;
    return L;
}

SRLA_HD inline Front16Layout make_front16_layout(uint32_t nmax, uint32_t fft_max, uint32_t nch)
{
    Front16Layout L;
    const uint32_t n8 = round_up_u32(nmax, 8);
    uint32_t off = 0;
    L.region_off = off; off += round_up_u32(8u * fft_max, 16);
    L.row_bytes = round_up_u32(2u * n8 + 48u, 16);
    L.rows_off = off; off += nch * L.row_bytes;
    L.red_off = off; off += 8u * 4u * 10u;                  /* four warps x (eight int64 sums + two flags) */
    L.coef_off = off; off += 16u * 4u;
    L.bar_off = off; off += 16u;
    L.total = off;
    return L;
}

/* lpc kernels: lags and one coefficient vector of 32 candidates, interleaved [index][lane] */
struct LpcLayout { uint32_t total; uint32_t select_total; };
SRLA_HD inline LpcLayout make_lpc_layout(uint32_t P)
{
    LpcLayout L;
    L.total = 8u * 32u * ((P + 2u) + (P + 3u));
    L.select_total = 8u * 32u * (P + 3u) + (8u + 4u) * 4u * 32u;      /* coefficient vectors + per-warp best (bits, order) */
    return L;
}

/* residual_kernel: signal, residual + (packed 16-bit sample pairs of the FIR | mean pyramid), coefficients */
struct ResidLayout {
    uint32_t region_off, region_bytes; /* int32 residual[n4], then the FIR's packed pairs, later the mean pyramid (doubles) */
    uint32_t sig_off;
    uint32_t coef_off;                 /* int32 x (roundup4(P) + 4)                            */
    uint32_t coefb_off;                /* the same coefficients as int8 x 4 words              */
    uint32_t red_off;                  /* 1024 bytes of reduction scratch                      */
    uint32_t total;
};
/* front padding of the FIR's inputs: the taps of the first outputs after the warm-up reach up to roundup4(P)
 * samples in front of the block; those taps carry zero coefficients, so the padding only has to be addressable */
SRLA_HD inline uint32_t resid_front_pad(uint32_t P) { return round_up_u32(P, 4) + 4u; }             /* samples, multiple of 4 */
SRLA_HD inline uint32_t resid_pair_front(uint32_t P) { return round_up_u32(round_up_u32(P, 4) / 4u + 2u, 16); }   /* pair entries */

/* svr_kernel: the normalised signal, the residual of the current iterate (first used as int32 scratch of the
 * signal rebuild), seven vectors of P+1 doubles, scalars */
struct SvrLayout { uint32_t data_off, resid_off, vec_off, total; };
SRLA_HD inline SvrLayout make_svr_layout(uint32_t nmax, uint32_t P)
{
    SvrLayout L;
    const uint32_t n4 = round_up_u32(nmax, 4);
    uint32_t off = 0;
    L.data_off = off; off += 8u * (n4 + 32u);
    L.resid_off = off; off += 8u * n4 + 128u;
    L.vec_off = off; off += 8u * 7u * (P + 1u) + 64u;
    L.total = off;
    return L;
}

SRLA_HD inline ResidLayout make_resid_layout(uint32_t nmax, uint32_t P)
{
    ResidLayout L;
    const uint32_t n4 = round_up_u32(nmax, 4);
    const uint32_t parts = (nmax < (uint32_t)kMaxParts) ? nmax : (uint32_t)kMaxParts;
    const uint32_t pyramid = 16u * round_up_u32(parts, 2) + 16u;
    const uint32_t pairs = 4u * round_up_u32(nmax, 8) + 32u + 16u * resid_pair_front(P);      /* one 16-byte entry per 4 samples */
    const uint32_t scratch = (pyramid > pairs) ? pyramid : pairs;
    uint32_t off = 0;
    L.region_off = off;
    L.region_bytes = round_up_u32(4u * n4 + (scratch > 8192u ? scratch : 8192u), 16);     /* the Rice search parks 64 + 2048 + 5632 bytes there */
    off += L.region_bytes;
    L.sig_off = off; off += 4u * (n4 + 12u + resid_front_pad(P));
    L.coef_off = off; off += 4u * (round_up_u32(P, 4) + 4u);
    L.coefb_off = off; off += 4u * (round_up_u32(P, 4) / 4u + 4u);
    L.red_off = off; off += 1024u;
    L.total = off;
    return L;
}

/* residual16_kernel (persistent CTAs, 16-bit PCM staged by bulk asynchronous copies): two buffers that hold the int16 source
 * rows of the NEXT candidate while the int32 residual of the current one lives where its own rows were, the pair entries /
 * int32 signal / mean pyramid scratch, coefficients, reduction scratch, the two item descriptors and their mbarriers */
struct Resid16Layout {
    uint32_t buf_off[2], buf_bytes, row_bytes;
    uint32_t scratch_off, coef_off, coefb_off, red_off, desc_off, stage_off, bar_off, total;
};
SRLA_HD inline Resid16Layout make_resid16_layout(uint32_t nmax, uint32_t P)
{
    Resid16Layout L;
    const uint32_t n8 = round_up_u32(nmax, 8);
    const uint32_t parts = (nmax < (uint32_t)kMaxParts) ? nmax : (uint32_t)kMaxParts;
    const uint32_t pyramid = 16u * round_up_u32(parts, 2) + 16u;
    const uint32_t pairs = 4u * n8 + 32u + 16u * resid_pair_front(P);
    const uint32_t sigbytes = 4u * (n8 + 12u + resid_front_pad(P));
    uint32_t scratch = (pyramid > pairs) ? pyramid : pairs;
    if (sigbytes > scratch) { scratch = sigbytes; }
    if (scratch < 8192u) { scratch = 8192u; }
    L.row_bytes = round_up_u32(2u * n8 + 16u, 16);
    L.buf_bytes = round_up_u32((4u * n8 > 2u * L.row_bytes) ? 4u * n8 : 2u * L.row_bytes, 128);
    uint32_t off = 0;
    L.buf_off[0] = off; off += L.buf_bytes;
    L.buf_off[1] = off; off += L.buf_bytes;
    L.scratch_off = off; off += round_up_u32(scratch, 16);
    L.coef_off = off; off += 4u * (round_up_u32(P, 4) + 4u);
    L.coefb_off = off; off += round_up_u32(4u * (round_up_u32(P, 4) / 4u + 4u), 16);
    L.red_off = off; off += 1024u;
    L.desc_off = off;
    L.stage_off = off; off += 2u * 128u;           /* two slots: job head (16) + candidate record head (64) + stream (48), fetched with cp.async */
    L.bar_off = off; off += 16u;
    L.total = off;
    return L;
}

/* front_big_kernel (blocks beyond kMaxSharedBlock): everything a job needs, in global memory per CTA */
struct FrontBigLayout { uint32_t raw_off, sig_off, a_off, b_off, pbuf_off, lags_off, total; };
SRLA_HD inline FrontBigLayout make_front_big_layout(uint32_t nmax, uint32_t fft_max)
{
    FrontBigLayout L;
    const uint32_t n4 = round_up_u32(nmax, 4);
    uint32_t off = 0;
    L.raw_off = off; off += 4u * (n4 + 32u);
    L.sig_off = off; off += 4u * (n4 + 32u);
    L.a_off = off; off += 8u * fft_max;                 /* fft_max / 2 complex elements */
    L.b_off = off; off += 8u * fft_max;
    L.pbuf_off = off; off += 8u * (fft_max + 272u);
    L.lags_off = off; off += 8u * 272u;
    L.total = round_up_u32(off, 256);
    return L;
}

} // namespace srla
#endif
