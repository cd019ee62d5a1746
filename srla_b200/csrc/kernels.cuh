/*
 * kernels.cuh -- the sm_100a CUDA kernels of the SRLA encode path.
 *
 *   lshift_jobs_kernel / lshift_finish_kernel whole-stream OR reduce -> trailing-zero shift   (A0)
 *   front_kernel<BPT>       one CTA per (block, candidate channel): mid/side, pre-emphasis, optional LTP,
 *                           Welch window + FFT autocorrelation, block resident in shared memory (A2-A5, A11)
 *   lpc_levinson_kernel     one thread per candidate: Levinson-Durbin, reflection coefficients + error variances (A5)
 *   lpc_select_kernel       order choice (parallel over orders), coefficient rebuild, quantisation          (A6)
 *   residual_kernel         one CTA per candidate: int32 FIR residual, Rice search, side-info bits   (A3, A7, A8)
 *   decide_kernel           block type, stereo method, exact block size                       (A1, A2)
 *   scan_kernel             output offsets of the blocks / streams
 *   emit_kernel             one CTA per block: header, side information, Rice codes, Fletcher (A1, A9, A10)
 *
 * Floating point: this translation unit MUST be compiled with -fmad=false.  The reference is ISO C90
 * (no FMA contraction); byte-identical output needs every double operation rounded separately and
 * in the reference's order.  Division and sqrt are IEEE (nvcc defaults -prec-div/-prec-sqrt=true).
 */
#ifndef SRLA_B200_KERNELS_CUH
#define SRLA_B200_KERNELS_CUH

#include <cfloat>
#include <cuda_runtime.h>

#include "types.h"

namespace srla {

/* ------------------------------------------------------------------------------------------------
 * small helpers
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t zigzag32(int32_t v) { return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31); }   /* srla_utility.h:31 */
__device__ __forceinline__ int32_t asr32(int32_t v, uint32_t s) { return (s >= 32u) ? (v >> 31) : (v >> s); }
__device__ __forceinline__ double round_half_away(double d) { return (d >= 0.0) ? floor(d + 0.5) : -floor(-d + 0.5); } /* srla_utility.c:22 */
__device__ __forceinline__ uint32_t ceil_pow2_u32(uint32_t v) { return (v <= 1u) ? 1u : (1u << (32 - __clz(v - 1u))); }

__device__ __forceinline__ int32_t load_sample(const StreamDev &st, uint32_t ch, uint32_t idx)
{
    const unsigned long long at = (unsigned long long)ch * st.stride + idx;
    return (st.sample_bytes == 2u) ? (int32_t)__ldg(reinterpret_cast<const short *>(st.pcm) + at)
                                   : __ldg(reinterpret_cast<const int32_t *>(st.pcm) + at);
}

/* PCM is streamed once per candidate: keep it out of L1 so the FFT twiddle tables stay resident there */
__device__ __forceinline__ int2 ldg_stream_v2(const void *ptr)
{
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ptr));
    return v;
}
__device__ __forceinline__ int4 ldg_stream_v4(const void *ptr)
{
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
    return v;
}
__device__ __forceinline__ int4 load_quad(const StreamDev &st, uint32_t ch, uint32_t idx)
{
    const unsigned long long at = (unsigned long long)ch * st.stride + idx;
    if (st.sample_bytes == 2u) {
        const int2 v = ldg_stream_v2(reinterpret_cast<const short *>(st.pcm) + at);
        return make_int4((int32_t)(short)(v.x & 0xffff), v.x >> 16, (int32_t)(short)(v.y & 0xffff), v.y >> 16);
    }
    return ldg_stream_v4(reinterpret_cast<const int32_t *>(st.pcm) + at);
}
__device__ __forceinline__ bool quad_aligned(const StreamDev &st, uint32_t ch, uint32_t idx)
{
    const unsigned long long addr = reinterpret_cast<unsigned long long>(st.pcm) + ((unsigned long long)ch * st.stride + idx) * st.sample_bytes;
    return (addr & (4ull * st.sample_bytes - 1ull)) == 0ull;
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
    return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
    return v;
}

/* inclusive block scan of one uint32 per thread; returns the inclusive prefix, *total = block sum.
 * scratch: kWarps + 1 words */
__device__ __forceinline__ uint32_t block_scan_inclusive(uint32_t v, uint32_t *scratch, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) { x += y; } }
    __syncthreads();                       /* scratch may still be read from a previous call */
    if (lane == 31) { scratch[warp] = x; }
    __syncthreads();
    uint32_t base = 0, sum = 0;
    #pragma unroll
    for (int w = 0; w < kWarps; ++w) { const uint32_t s = scratch[w]; if (w < warp) { base += s; } sum += s; }
    *total = sum;
    return x + base;
}

/* ------------------------------------------------------------------------------------------------
 * A0: trailing-zero shift of a whole stream (srla_utility.c:177-203).
 * lshift_jobs_kernel: one CTA per job ORs the job's samples of every channel into its stream's
 * or_mask (16-byte vector loads where aligned) -- the only purely HBM-bound kernel of the path.
 * lshift_finish_kernel: or_mask -> trailing zero count; optionally snapshots the value the jobs of
 * this launch group are about to use (pipelined host path, see Runner::run).
 * ---------------------------------------------------------------------------------------------- */
/* 16-byte loads of 16-bit PCM (eight samples each), several per thread in flight: the pass that only ORs samples is bound
 * by HBM and by how many bytes a CTA keeps in flight */
__device__ __forceinline__ uint32_t or_row_int16(const short *row, uint32_t n, uint32_t tid, uint32_t nthreads)
{
    uint32_t acc = 0;
    const uint32_t nvec = n >> 3;
    const int4 *v = reinterpret_cast<const int4 *>(row);
    uint32_t g = tid;
    for (; g + 3u * nthreads < nvec; g += 4u * nthreads) {
        const int4 a = ldg_stream_v4(v + g), b = ldg_stream_v4(v + g + nthreads), c = ldg_stream_v4(v + g + 2u * nthreads), d = ldg_stream_v4(v + g + 3u * nthreads);
        acc |= (uint32_t)(a.x | a.y | a.z | a.w) | (uint32_t)(b.x | b.y | b.z | b.w) | (uint32_t)(c.x | c.y | c.z | c.w) | (uint32_t)(d.x | d.y | d.z | d.w);
    }
    for (; g < nvec; g += nthreads) { const int4 a = ldg_stream_v4(v + g); acc |= (uint32_t)(a.x | a.y | a.z | a.w); }
    for (uint32_t i = (nvec << 3) + tid; i < n; i += nthreads) { acc |= (uint32_t)(int32_t)__ldg(row + i); }
    /* only the lowest set bit of the mask matters (trailing-zero count) and it lies in the 16 raw bits of either half */
    return (acc & 0xffffu) | (acc >> 16);
}

__global__ void __launch_bounds__(256) lshift_jobs_kernel(StreamDev *streams, const Job *jobs, uint32_t nch, uint32_t num_jobs)
{
  for (uint32_t jb = blockIdx.x; jb < num_jobs; jb += gridDim.x) {
    const Job &job = jobs[jb];
    StreamDev &st = streams[job.stream];
    const uint32_t n = job.nsmpl;
    uint32_t acc = 0;
    for (uint32_t c = 0; c < nch; ++c) {
        if (st.sample_bytes == 2u) {
            const short *row = reinterpret_cast<const short *>(st.pcm) + (unsigned long long)c * st.stride + job.offset;
            if ((reinterpret_cast<unsigned long long>(row) & 15ull) == 0ull) { acc |= or_row_int16(row, n, threadIdx.x, blockDim.x); continue; }
        }
        const bool vec = quad_aligned(st, c, job.offset);
        const uint32_t nquad = vec ? (n >> 2) : 0u;
        for (uint32_t g = threadIdx.x; g < nquad; g += blockDim.x) {
            const int4 q = load_quad(st, c, job.offset + 4u * g);
            acc |= (uint32_t)(q.x | q.y | q.z | q.w);
        }
        for (uint32_t i = 4u * nquad + threadIdx.x; i < n; i += blockDim.x) { acc |= (uint32_t)load_sample(st, c, job.offset + i); }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc |= __shfl_xor_sync(0xffffffffu, acc, o); }
    /* the mask only ever gains bits: skip the atomic when this warp adds nothing new (almost always after the first blocks) */
    if ((threadIdx.x & 31) == 0 && (acc & ~*reinterpret_cast<volatile uint32_t *>(&st.or_mask))) { atomicOr(&st.or_mask, acc); }
  }
}

/* ------------------------------------------------------------------------------------------------
 * WAV ingest.  The reference CLI reads a WAV data chunk one sample at a time through a bit buffer
 * and converts it to planar sign-extended int32 (libs/wav/src/wav.c:543-553, conversions :841-866: 8-bit is
 * unsigned with offset 128, 16- and 24-bit are little-endian two's complement).  Here the data chunk is copied
 * to HBM as it lies in the file and one CTA per job writes the planar layout the analysis kernels read
 * (int16 for sources of at most 16 bits, else int32).  The samples are OR-ed into the stream's or_mask on the
 * way (srla_utility.c:177-203), so no lshift_jobs_kernel pass is needed over data that came in this way.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) deinterleave_jobs_kernel(StreamDev *streams, const Job *jobs, uint32_t nch)
{
    const Job &job = jobs[blockIdx.x];
    StreamDev &st = streams[job.stream];
    const uint32_t n = job.nsmpl, cb = st.container_bytes;
    const unsigned char *src = reinterpret_cast<const unsigned char *>(st.raw) + (size_t)job.offset * nch * cb;
    uint32_t acc = 0;
    uint32_t done = 0;                                   /* frames handled by the vector path */
    if (cb == 2u && nch == 2u && st.sample_bytes == 2u && (reinterpret_cast<uintptr_t>(src) & 15u) == 0u
        && quad_aligned(st, 0, job.offset) && quad_aligned(st, 1, job.offset)) {
        /* 16-bit stereo: four frames per 16-byte load, one 8-byte store per channel */
        short *left = reinterpret_cast<short *>(const_cast<void *>(st.pcm)) + job.offset;
        short *right = left + st.stride;
        const uint32_t nquad = n >> 2;
        for (uint32_t g = threadIdx.x; g < nquad; g += blockDim.x) {
            const int4 v = ldg_stream_v4(src + 16u * (size_t)g);
            int2 l, r;
            l.x = (int32_t)__byte_perm((uint32_t)v.x, (uint32_t)v.y, 0x5410); r.x = (int32_t)__byte_perm((uint32_t)v.x, (uint32_t)v.y, 0x7632);
            l.y = (int32_t)__byte_perm((uint32_t)v.z, (uint32_t)v.w, 0x5410); r.y = (int32_t)__byte_perm((uint32_t)v.z, (uint32_t)v.w, 0x7632);
            /* only the lowest set bit of the mask matters (trailing-zero count) and it lies in the 16 raw bits */
            const uint32_t any = (uint32_t)(v.x | v.y | v.z | v.w);
            acc |= (any & 0xffffu) | (any >> 16);
            *reinterpret_cast<int2 *>(left + 4u * g) = l;
            *reinterpret_cast<int2 *>(right + 4u * g) = r;
        }
        done = nquad << 2;
    }
    const uint32_t total = (n - done) * nch;
    const unsigned char *tail = src + (size_t)done * nch * cb;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        const uint32_t frame = i / nch, ch = i - frame * nch;
        const unsigned char *q = tail + (size_t)i * cb;
        int32_t v;
        if (cb == 1u) { v = (int32_t)q[0] - 128; }
        else if (cb == 2u) { v = (int32_t)(short)((uint32_t)q[0] | ((uint32_t)q[1] << 8)); }
        else { v = (int32_t)(((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16)) << 8) >> 8; }
        acc |= (uint32_t)v;
        const size_t at = (size_t)ch * st.stride + job.offset + done + frame;
        if (st.sample_bytes == 2u) { reinterpret_cast<short *>(const_cast<void *>(st.pcm))[at] = (short)v; }
        else { reinterpret_cast<int32_t *>(const_cast<void *>(st.pcm))[at] = v; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc |= __shfl_xor_sync(0xffffffffu, acc, o); }
    if ((threadIdx.x & 31) == 0 && acc) { atomicOr(&st.or_mask, acc); }
}

__global__ void lshift_finish_kernel(StreamDev *streams, uint32_t num_streams, uint32_t *snapshot)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < num_streams) {
        const uint32_t m = streams[s].or_mask;
        const uint32_t sh = m ? (uint32_t)(__ffs((int)m) - 1) : 0u;
        streams[s].lshift = sh;
        if (snapshot) { snapshot[s] = sh; }
    }
}

/* ------------------------------------------------------------------------------------------------
 * FFT exactly as the reference evaluates it (libs/fft/src/fft.c:71-128, 147-198).
 *
 * The reference's radix-4 Stockham stages are executed IN PLACE in shared memory: every thread
 * first pulls all inputs of its work unit into registers, the CTA synchronises, then the outputs
 * are stored.  TWO consecutive radix-4 stages are fused into one pass over shared memory (a work
 * unit = 16 points = 4 butterflies of stage t feeding 4 butterflies of stage t+1 in registers; the
 * trailing radix-4 + radix-2 pair is an 8-point unit), which halves the shared-memory traffic that
 * bounds this kernel.  Every output is the same expression tree as in the reference, with the
 * host-tabulated twiddle recurrence.  Complex element i lives at slot fft_slot(i): an XOR swizzle
 * that keeps both the unit-strided stores of the first pass and all contiguous accesses free of
 * bank conflicts.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ uint32_t fft_slot(uint32_t i) { return i ^ ((i >> 4) & 7u); }

struct Twiddle3 { double2 w1, w2, w3; };
/* table of one stage size ns: w1[ns/4], w2[ns/4], w3[ns/4] (host_tables.h) */
__device__ __forceinline__ Twiddle3 load_twiddle(const double2 *table, uint32_t quarter, uint32_t p)
{
    Twiddle3 t;
    t.w1 = __ldg(table + p); t.w2 = __ldg(table + quarter + p); t.w3 = __ldg(table + 2u * quarter + p);
    return t;
}

/* one radix-4 butterfly of the FORWARD transform, fft.c:93-105.
 * The inverse transform (flag = +1: conjugate twiddles, opposite rotation) is evaluated as
 * conj(forward(conj(x))): IEEE round-to-nearest is sign-symmetric, so every intermediate of that
 * evaluation is the exact conjugate of the reference's and the results are bit-identical -- and the
 * kernel carries one copy of the transform code instead of two (it is bound by instruction fetch). */
__device__ __forceinline__ void butterfly4(const double2 a, const double2 b, const double2 c, const double2 d, const Twiddle3 &w,
                                           double2 &y0, double2 &y1, double2 &y2, double2 &y3)
{
    const double2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), bmd = csub(b, d);
    /* j * (b - d), j = (0, +1) -> (-im, re)   (the reference's 0.0 * x terms only affect the sign of zeros) */
    const double2 jbmd = make_double2(-bmd.y, bmd.x);
    y0 = cadd(apc, bpd);
    y1 = cmul(w.w1, csub(amc, jbmd));
    y2 = cmul(w.w2, csub(apc, bpd));
    y3 = cmul(w.w3, cadd(amc, jbmd));
}

/* butterfly whose twiddles are the first entries of a stage table, exactly (1, 0): multiplying by them returns the
 * operand (up to the sign of a zero), so the three complex products are skipped */
__device__ __forceinline__ void butterfly4_unit(const double2 a, const double2 b, const double2 c, const double2 d,
                                                double2 &y0, double2 &y1, double2 &y2, double2 &y3)
{
    const double2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), bmd = csub(b, d);
    const double2 jbmd = make_double2(-bmd.y, bmd.x);
    y0 = cadd(apc, bpd);
    y1 = csub(amc, jbmd);
    y2 = csub(apc, bpd);
    y3 = cadd(amc, jbmd);
}

/* exact int32 -> double without the quarter-rate I2F.F64: 2^52 + 2^31 + x is representable */
__device__ __forceinline__ double int_to_double(int32_t x)
{
    return __hiloint2double(0x43300000, (int)((uint32_t)x ^ 0x80000000u)) - 4503601774854144.0;
}

/* Welch-windowed input of the autocorrelation (lpc.c:252-266, srla_encoder.c:1061-1064): complex element e of
 * the packed real transform = samples (2e, 2e+1), zero beyond n.  kPre: sig holds the UNFILTERED candidate
 * and the pre-emphasis (srla_utility.c:342-358, filter memory = first sample) is applied on the fly. */
/* kChain (front_tail_kernel only): for an odd n the reference's window loop never writes the middle sample of its
 * scratch buffer (lpc.c:260-264), which therefore still holds what the previous call's inverse transform left at that
 * index: `stale` */
template <bool kPre, bool kChain = false>
struct WindowSource {
    /* full blocks of 8192 samples in a CTA of 256 threads run the literal-size passes (see welch_autocorr_core) */
    static constexpr uint32_t kFixedM = kChain ? 0u : 4096u, kFixedThreads = kChain ? 0u : 256u;
    __device__ __forceinline__ void first_pass_fixed(double2 *x, const double2 *tw_a, const double2 *tw_b) const;
    const int32_t *sig; uint32_t n, half_n, pc; double unit, div, dn1;
    double stale;
    bool full;                    /* n is the transform size: no zero padding, the window halves meet at n / 2 */
    /* element whose two samples are 2e, 2e+1 with window arguments ds0, ds0 + step (exact doubles supplied by the caller) */
    __device__ __forceinline__ double2 element_at(uint32_t e, double ds0, double step) const
    {
        const uint32_t i = 2u * e;
        const int2 c = *reinterpret_cast<const int2 *>(sig + i);
        int32_t x0 = c.x, x1 = c.y;
        if (kPre) {
            const int32_t prv = sig[i ? i - 1u : 0u];
            x1 = (int32_t)((uint32_t)c.y - (uint32_t)((int32_t)((uint32_t)c.x * pc) >> 4));
            x0 = (int32_t)((uint32_t)c.x - (uint32_t)((int32_t)((uint32_t)prv * pc) >> 4));
        }
        const double ds1 = ds0 + step;
        const double w0 = div * ds0 * (dn1 - ds0), w1 = div * ds1 * (dn1 - ds1);
        return make_double2(int_to_double(x0) * w0, int_to_double(x1) * w1);
    }
    __device__ __forceinline__ double one(uint32_t i, int32_t cur, int32_t prv) const
    {
        if (i >= n) { return 0.0; }
        if (kChain && (n & 1u) && i == half_n) { return stale; }
        const int32_t x = kPre ? (int32_t)((uint32_t)cur - (uint32_t)((int32_t)((uint32_t)prv * pc) >> 4)) : cur;
        const uint32_t s = (i < half_n) ? i : (n - 1u - i);
        const double ds = int_to_double((int32_t)s);
        const double w = div * ds * (dn1 - ds);              /* (n - 1 - s) as an exact double difference */
        return int_to_double(x) * w;
    }
    __device__ __forceinline__ double2 element(uint32_t e) const
    {
        const uint32_t i = 2u * e;
        if (i >= n) { return make_double2(0.0, 0.0); }
        /* sig is zero-padded past n and has 4 readable samples in front of index 0 */
        const int2 c = *reinterpret_cast<const int2 *>(sig + i);
        const int32_t prv = kPre ? sig[i ? i - 1u : 0u] : 0;
        return make_double2(one(i, c.x, prv), one(i + 1u, c.y, c.x));
    }
    /* the first pass of the forward transform, which windows the samples itself (defined behind fft_pair_pass) */
    __device__ __forceinline__ void first_pass(double2 *x, uint32_t M, uint32_t nn, uint32_t lgs, const double2 *tw_a, const double2 *tw_b, uint32_t need) const;
    /* the 16 inputs of thread tid's work unit in the first pass of an M-point transform WITHOUT padding (full): elements
     * with j < 2 lie in the rising window half (argument s = i), the others in the falling half (s = n - 1 - i); all
     * arguments are exact doubles derived from two conversions per thread */
    __device__ __forceinline__ void load_full(double2 (&v)[4][4], const uint32_t tid, const uint32_t M) const
    {
        const double lo0 = int_to_double((int32_t)(2u * tid)), hi0 = int_to_double((int32_t)(n - 1u - 2u * tid));
        const double dq = int_to_double((int32_t)(M >> 3));            /* samples between consecutive jp */
        #pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t e = tid + (uint32_t)jp * (M >> 4) + (uint32_t)j * (M >> 2);
                const double off = dq * (double)(jp + 4 * j);          /* 2 * (e - tid), exact */
                v[jp][j] = (j < 2) ? element_at(e, lo0 + off, 1.0) : element_at(e, hi0 - off, -1.0);
            }
        }
    }
};

/* complex FFT of M points (M a power of two, M/16 <= blockDim.x, M/8 <= blockDim.x when M = 8 * 4^k).
 * `need`: only outputs 0..need-1 are required (M = all).  The autocorrelation reads just the first lags
 * of the inverse transform, so its last two passes skip every butterfly none of whose outputs can reach
 * them -- the surviving outputs are the reference's expression trees unchanged.
 * `src` (forward transform, M >= 16 only): the first pass takes its inputs from the windowed samples
 * instead of x, which saves one full write + read of the buffer. */
/* one fused pair of radix-4 stages (sizes nn and nn/4, output stride 1 << lgs).
 * `need`: only outputs 0..need-1 of the whole transform are required (M = all); in the last pair pass an
 * output at position o is read later only when (o mod (16 << lgs)) < need, so butterflies that cannot
 * reach a required output are skipped -- the surviving outputs are the reference's expression trees. */
/* stands in for the sample source of the passes that read the buffer */
struct NoSource {
    static constexpr bool full = false;
    __device__ __forceinline__ void load_full(double2 (&)[4][4], uint32_t, uint32_t) const {}
    __device__ __forceinline__ double2 element(uint32_t) const { return make_double2(0.0, 0.0); }
};

template <typename Src, bool kFromSamples, bool kAllActive = false>       /* kAllActive: the CTA has exactly M / 16 threads */
__device__ __forceinline__ void fft_pair_pass_impl(double2 *x, const uint32_t M, const uint32_t nn, const uint32_t lgs,
                                                   const double2 *tw_a, const double2 *tw_b, const uint32_t need, const Src &src)
{
    /* called with literal M, nn, lgs (the *_fixed wrappers below) everything but the data path folds away */
    const uint32_t tid = threadIdx.x;
    const uint32_t units = M >> 4;
    const bool last_pair = (nn < 256u);          /* outputs feed the tail pass (or are final) */
    const bool active = kAllActive || tid < units;
    const uint32_t q = tid & ((1u << lgs) - 1u), p0 = tid >> lgs;
    const bool fast_load = ((M >> 4) & 127u) == 0u;
    double2 v[4][4];
    if (active) {
        if (kFromSamples && src.full) {
            src.load_full(v, tid, M);                                      /* no padding: the window halves meet at M / 2 */
        } else if (!kFromSamples && fast_load) {
            /* M / 16 is a multiple of 128: the swizzle term (e >> 4) & 7 is the same for all 16 inputs of the unit */
            const double2 *b = x + (tid ^ ((tid >> 4) & 7u));
            #pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                #pragma unroll
                for (int j = 0; j < 4; ++j) { v[jp][j] = b[(uint32_t)(jp + 4 * j) * (M >> 4)]; }
            }
        } else {
            #pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t e = tid + (uint32_t)jp * (M >> 4) + (uint32_t)j * (M >> 2);
                    v[jp][j] = kFromSamples ? src.element(e) : x[fft_slot(e)];
                }
            }
        }
    }
    (void)fast_load;
    if (!kFromSamples) { __syncthreads(); }      /* in place: every input is in registers before any output is stored */
    if (active) {
        double2 y[4][4];
        #pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            const Twiddle3 wa = load_twiddle(tw_a, nn >> 2, p0 + (uint32_t)jp * (nn >> 4));
            butterfly4(v[jp][0], v[jp][1], v[jp][2], v[jp][3], wa, y[jp][0], y[jp][1], y[jp][2], y[jp][3]);
        }
        const Twiddle3 wb = load_twiddle(tw_b, nn >> 4, p0);
        if (lgs == 0u && (!last_pair || need >= M)) {
            /* outputs 16 tid + (j + 4 i): the swizzle term is tid & 7 */
            double2 *b = x + 16u * tid;
            const uint32_t t = tid & 7u;
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                double2 z0, z1, z2, z3;
                butterfly4(y[0][j], y[1][j], y[2][j], y[3][j], wb, z0, z1, z2, z3);
                b[(uint32_t)j ^ t] = z0; b[(uint32_t)(j + 4) ^ t] = z1; b[(uint32_t)(j + 8) ^ t] = z2; b[(uint32_t)(j + 12) ^ t] = z3;
            }
        } else if (lgs == 4u && (!last_pair || need >= M)) {
            /* outputs q + 16 j + 64 i + 256 p0: the swizzle term is (j + 4 i) & 7, a constant per output */
            double2 *b = x + 256u * p0;
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                double2 z0, z1, z2, z3;
                butterfly4(y[0][j], y[1][j], y[2][j], y[3][j], wb, z0, z1, z2, z3);
                b[(q ^ (uint32_t)(j & 7)) + 16u * j]              = z0;
                b[(q ^ (uint32_t)((j + 4) & 7)) + 16u * j + 64u]  = z1;
                b[(q ^ (uint32_t)(j & 7)) + 16u * j + 128u]       = z2;
                b[(q ^ (uint32_t)((j + 4) & 7)) + 16u * j + 192u] = z3;
            }
        } else if (!last_pair || need >= M) {
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                double2 z0, z1, z2, z3;
                butterfly4(y[0][j], y[1][j], y[2][j], y[3][j], wb, z0, z1, z2, z3);
                const uint32_t o = q + ((uint32_t)j << lgs) + ((16u * p0) << lgs);
                x[fft_slot(o)]                = z0;
                x[fft_slot(o + (4u << lgs))]  = z1;
                x[fft_slot(o + (8u << lgs))]  = z2;
                x[fft_slot(o + (12u << lgs))] = z3;
            }
        } else {
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t low = q + ((uint32_t)j << lgs);
                const uint32_t o = low + ((16u * p0) << lgs);
                if (low + (4u << lgs) < need) {
                    double2 z0, z1, z2, z3;
                    butterfly4(y[0][j], y[1][j], y[2][j], y[3][j], wb, z0, z1, z2, z3);
                    x[fft_slot(o)]                = z0;
                    x[fft_slot(o + (4u << lgs))]  = z1;
                    x[fft_slot(o + (8u << lgs))]  = z2;
                    x[fft_slot(o + (12u << lgs))] = z3;
                } else if (low < need) {
                    x[fft_slot(o)] = cadd(cadd(y[0][j], y[2][j]), cadd(y[1][j], y[3][j]));     /* y0 of butterfly4 */
                }
            }
        }
    }
    __syncthreads();
}

/* a real call, not inlined: each pass gets its own register allocation and the kernel one copy of it */
template <typename Src, bool kFromSamples>
__device__ __noinline__ void fft_pair_pass(double2 *x, const uint32_t M, const uint32_t nn, const uint32_t lgs,
                                           const double2 *tw_a, const double2 *tw_b, const uint32_t need, const Src src)
{
    fft_pair_pass_impl<Src, kFromSamples>(x, M, nn, lgs, tw_a, tw_b, need, src);
}

/* the same pass with the transform size, stage size and stride known at compile time (kNeedAll: every output is needed) */
template <uint32_t kM, uint32_t kNN, uint32_t kLgs, bool kNeedAll>
__device__ __noinline__ void fft_pair_pass_fixed(double2 *x, const double2 *tw_a, const double2 *tw_b, const uint32_t need)
{
    fft_pair_pass_impl<NoSource, false, true>(x, kM, kNN, kLgs, tw_a, tw_b, kNeedAll ? kM : need, NoSource());
}

template <bool kPre, bool kChain>
__device__ __forceinline__ void WindowSource<kPre, kChain>::first_pass(double2 *x, uint32_t M, uint32_t nn, uint32_t lgs, const double2 *tw_a, const double2 *tw_b, uint32_t need) const
{
    fft_pair_pass<WindowSource<kPre, kChain>, true>(x, M, nn, lgs, tw_a, tw_b, need, *this);
}

/* WindowSource's first pass with the transform size a literal (a full block: n == 2 kM); scalar arguments travel in registers */
template <bool kPre, uint32_t kM>
__device__ __noinline__ void fft_first_pass_sig_fixed(double2 *x, const double2 *tw_a, const double2 *tw_b, const uint32_t sig_shared, const uint32_t pc, const double div)
{
    WindowSource<kPre, false> ws;
    ws.sig = reinterpret_cast<const int32_t *>(__cvta_shared_to_generic((size_t)sig_shared));
    ws.n = 2u * kM; ws.half_n = kM; ws.pc = pc; ws.unit = 1.0; ws.div = div; ws.dn1 = (double)(int32_t)(2u * kM - 1u); ws.stale = 0.0; ws.full = true;
    fft_pair_pass_impl<WindowSource<kPre, false>, true, true>(x, kM, kM, 0u, tw_a, tw_b, kM, ws);
}
template <bool kPre, bool kChain>
__device__ __forceinline__ void WindowSource<kPre, kChain>::first_pass_fixed(double2 *x, const double2 *tw_a, const double2 *tw_b) const
{
    if constexpr (!kChain) { fft_first_pass_sig_fixed<kPre, 4096u>(x, tw_a, tw_b, (uint32_t)__cvta_generic_to_shared(sig), pc, div); }
}

/* the stages left after the fused pairs: nn = 8 (radix-4 + radix-2), 4 (radix-4) or 2 (radix-2).  A tail unit
 * reads and writes exactly the same positions (u + k * s), so units are independent of each other: no barrier
 * between loads and stores, and a thread walks its units one at a time.  Units beyond `need` are skipped. */
__device__ __forceinline__ void fft_tail_pass_impl(double2 *x, const uint32_t M, const uint32_t nn, const double2 *tw8, const double2 *tw4, const uint32_t need, const uint32_t T)
{
    const uint32_t tid = threadIdx.x;
    if (nn == 8u) {
        /* radix-4 stage of size 8 (s = M/8) fused with the final radix-2 stage (fft.c:114-123) */
        const uint32_t s = M >> 3;
        const uint32_t live = (need < s) ? need : s;
        const Twiddle3 w1 = load_twiddle(tw8, 2u, 1u);             /* entry 0 is (1, 0): butterfly4_unit */
        #pragma unroll 1
        for (uint32_t u = tid; u < live; u += T) {
            double2 v[2][4], y[2][4];
            #pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                #pragma unroll
                for (int j = 0; j < 4; ++j) { v[pp][j] = x[fft_slot((uint32_t)pp * s + u + (uint32_t)j * (M >> 2))]; }
            }
            butterfly4_unit(v[0][0], v[0][1], v[0][2], v[0][3], y[0][0], y[0][1], y[0][2], y[0][3]);
            butterfly4(v[1][0], v[1][1], v[1][2], v[1][3], w1, y[1][0], y[1][1], y[1][2], y[1][3]);
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                x[fft_slot(u + (uint32_t)j * s)]            = cadd(y[0][j], y[1][j]);
                x[fft_slot(u + (uint32_t)j * s + (M >> 1))] = csub(y[0][j], y[1][j]);
            }
        }
    } else if (nn == 4u) {
        /* single radix-4 stage of size 4 (s = M/4): twiddle index 0 only */
        const uint32_t s = M >> 2;
        const uint32_t live = (need < s) ? need : s;
        (void)tw4;                                                  /* the only twiddle of this stage is (1, 0) */
        #pragma unroll 1
        for (uint32_t u = tid; u < live; u += T) {
            double2 y0, y1, y2, y3;
            butterfly4_unit(x[fft_slot(u)], x[fft_slot(u + s)], x[fft_slot(u + 2u * s)], x[fft_slot(u + 3u * s)], y0, y1, y2, y3);
            x[fft_slot(u)] = y0; x[fft_slot(u + s)] = y1; x[fft_slot(u + 2u * s)] = y2; x[fft_slot(u + 3u * s)] = y3;
        }
    } else if (nn == 2u) {
        const uint32_t s = M >> 1;
        const uint32_t live = (need < s) ? need : s;
        #pragma unroll 1
        for (uint32_t u = tid; u < live; u += T) {
            const double2 a = x[fft_slot(u)], b = x[fft_slot(u + s)];
            x[fft_slot(u)] = cadd(a, b); x[fft_slot(u + s)] = csub(a, b);
        }
    }
    __syncthreads();
}
__device__ __noinline__ void fft_tail_pass(double2 *x, const uint32_t M, const uint32_t nn, const double2 *tw8, const double2 *tw4, const uint32_t need)
{
    fft_tail_pass_impl(x, M, nn, tw8, tw4, need, blockDim.x);
}
template <uint32_t kM, uint32_t kNN, uint32_t kT, bool kNeedAll>
__device__ __noinline__ void fft_tail_pass_fixed(double2 *x, const double2 *tw8, const uint32_t need)
{
    fft_tail_pass_impl(x, kM, kNN, tw8, nullptr, kNeedAll ? kM : need, kT);
}

/* Welch window (lpc.c:252-266) + autocorrelation through the FFT (lpc.c:330-376).
 * sig[n] int32 in shared memory -> lags[0..nlags) (lags >= N read as 0.0).  buf: N doubles.
 * kPre: sig holds the UNFILTERED candidate and the pre-emphasis (srla_utility.c:342-358) is applied
 * on the fly with coefficient pre_coef. */
/* kChain (front_tail_kernel): `pbuf` mirrors the reference calculator's persistent scratch buffer in natural order
 * (lpc.c:211-213).  It supplies the stale middle sample of an odd-length window, receives the WHOLE inverse transform
 * of this call (no pruning), and the lags are read from it -- also those beyond the transform size, which the
 * reference copies from whatever earlier calls left there (lpc.c:371-373). */
/* forward split (fft.c:171-184), |X|^2 (lpc.c:355-362) and inverse split fused: all three only touch the element pair
 * (i, N/2 - i).  The inverse split's outputs are stored CONJUGATED. */
__device__ __forceinline__ void real_split_power(double2 *cx, const uint32_t N, const LaunchParams &p, const uint32_t tid, const uint32_t nthreads)
{
    const uint32_t M = N >> 1;
    const uint32_t lgN = 31u - (uint32_t)__clz((int)N);
    const double2 *tw = p.tw_real + p.tw_real_off[lgN];
    const uint32_t quarter = N >> 2;
    for (uint32_t i = 1u + tid; i <= quarter; i += nthreads) {
        const double2 w = __ldg(tw + (i - 1u));
        const double wr = w.x, wi_f = w.y, wi_b = -w.y;
        const uint32_t lo = i, hi = M - i;
        const double2 xl = cx[fft_slot(lo)], xh = cx[fft_slot(hi)];
        /* forward, flag = -1: c2 = -0.5 */
        double f1, f2, f3, f4;
        {
            const double c2 = -0.5;
            const double h1r = 0.5 * (xl.x + xh.x);
            const double h1i = 0.5 * (xl.y - xh.y);
            const double h2r = -c2 * (xl.y + xh.y);
            const double h2i = c2 * (xl.x - xh.x);
            f1 = h1r + (wr * h2r) - (wi_f * h2i);
            f2 = h1i + (wr * h2i) + (wi_f * h2r);
            f3 = h1r - (wr * h2r) + (wi_f * h2i);
            f4 = -h1i + (wr * h2i) + (wi_f * h2r);
        }
        /* for i == N/4 the pair is one element: the reference's second pair of stores wins */
        double p_lo, p_hi;
        if (lo == hi) { p_hi = f3 * f3 + f4 * f4; p_lo = p_hi; }
        else { p_lo = f1 * f1 + f2 * f2; p_hi = f3 * f3 + f4 * f4; }
        /* inverse, flag = +1: c2 = +0.5.  The power spectrum is real, so h1i and h2r of fft.c:171-184 are
         * exact zeros and every term they enter only adds +-0.0: g1 = h1r - wi*h2i, g2 = g4 = wr*h2i,
         * g3 = h1r + wi*h2i are the same values (up to the sign of a zero, which cannot reach a decision) */
        {
            const double h1r = 0.5 * (p_lo + p_hi);
            const double h2i = 0.5 * (p_lo - p_hi);
            const double t = wi_b * h2i, g2 = wr * h2i;
            if (lo != hi) { cx[fft_slot(lo)] = make_double2(h1r - t, -g2); }
            cx[fft_slot(hi)] = make_double2(h1r + t, -g2);
        }
    }
    if (tid == 0) {
        const double2 dc = cx[0];
        const double f0 = dc.x + dc.y, f1 = dc.x - dc.y;     /* forward DC / Nyquist */
        const double q0 = f0 * f0, q1 = f1 * f1;
        cx[0] = make_double2(0.5 * (q0 + q1), -(0.5 * (q0 - q1)));
    }
    __syncthreads();
}

/* the transform chain itself for a sample source `ws` (n >= 2) */
template <typename Src, bool kChain>
__device__ __forceinline__ void welch_autocorr_core(const Src &ws, const uint32_t n, double *buf, double *lags, const uint32_t lag_step, const uint32_t nlags,
                                                    const double ac_scale, const LaunchParams &p, double *pbuf);

template <bool kPre, bool kChain = false>
__device__ void welch_autocorr(const int32_t *sig, const int32_t pre_coef, const uint32_t n, double *buf, double *lags, const uint32_t lag_step, const uint32_t nlags,
                               const Job &job, const LaunchParams &p, double *pbuf = nullptr)
{
    const uint32_t tid = threadIdx.x, nthreads = blockDim.x;
    const uint32_t N = ceil_pow2_u32(n);
    const double unit = p.unit, div = job.welch_div;
    const uint32_t pc = (uint32_t)pre_coef;
    if (N < 2u) {
        if (tid == 0) {
            const double w = div * 0.0 * (double)(n - 1u);
            int32_t x = sig[0];
            if (kPre) { x = (int32_t)((uint32_t)x - (uint32_t)((int32_t)((uint32_t)x * pc) >> 4)); }    /* filter memory = first sample */
            buf[0] = ((double)x * unit) * w;
        }
        __syncthreads();
        for (uint32_t i = tid; i < nlags; i += nthreads) { lags[(size_t)i * lag_step] = (i < N) ? buf[i] * job.ac_scale : 0.0; }
        __syncthreads();
        return;
    }
    WindowSource<kPre, kChain> ws;
    ws.stale = kChain ? pbuf[n >> 1] : 0.0;
    /* unit = 2^-(bps-1) is a power of two: (x * unit) * w == x * (w * unit) exactly and w * unit == ((div * unit) * s) * (n-1-s)
     * exactly (no subnormals: |w| >= 4 / n^2 * 2^-23), so the scale rides on the divisor and costs no multiply per sample */
    ws.sig = sig; ws.n = n; ws.half_n = n >> 1; ws.pc = pc; ws.unit = 1.0; ws.div = div * unit; ws.dn1 = (double)(int32_t)(n - 1u);
    ws.full = (n == N) && (N >= 32u);
    welch_autocorr_core<WindowSource<kPre, kChain>, kChain>(ws, n, buf, lags, lag_step, nlags, job.ac_scale, p, pbuf);
}

template <typename Src, bool kChain>
__device__ __forceinline__ void welch_autocorr_core(const Src &ws, const uint32_t n, double *buf, double *lags, const uint32_t lag_step, const uint32_t nlags,
                                                    const double ac_scale, const LaunchParams &p, double *pbuf)
{
    const uint32_t tid = threadIdx.x, nthreads = blockDim.x;
    const uint32_t N = ceil_pow2_u32(n);
    double2 *cx = reinterpret_cast<double2 *>(buf);
    const uint32_t M = N >> 1;
    const uint32_t want = kChain ? N : ((nlags < N) ? nlags : N);
    if constexpr (Src::kFixedM != 0u) {
        /* the usual transform size of this source: the same passes with every size, stride and trip count a literal, so the
         * index arithmetic of the general code folds away (kernels that run with Src::kFixedThreads threads only) */
        static_assert(!kChain && ((Src::kFixedM == 2048u && Src::kFixedThreads == 128u) || (Src::kFixedM == 4096u && Src::kFixedThreads == 256u)),
                      "stage plans below: 2048 = 4^5 * 2 and 4096 = 4^6, one 16-point unit per thread");
        if (M == Src::kFixedM && ws.full && blockDim.x == Src::kFixedThreads) {                     /* a full block: n == 2 M */
            constexpr uint32_t FM = Src::kFixedM, FT = Src::kFixedThreads;
            const uint32_t need = (want + 1u) >> 1;
            if constexpr (FM == 2048u) {
                const double2 *t11 = p.tw_complex + p.tw_complex_off[11], *t9 = p.tw_complex + p.tw_complex_off[9];
                const double2 *t7 = p.tw_complex + p.tw_complex_off[7], *t5 = p.tw_complex + p.tw_complex_off[5], *t3 = p.tw_complex + p.tw_complex_off[3];
                ws.first_pass_fixed(cx, t11, t9);
                fft_pair_pass_fixed<FM, 128u, 4u, true>(cx, t7, t5, FM);
                fft_tail_pass_fixed<FM, 8u, FT, true>(cx, t3, FM);
                real_split_power(cx, 2u * FM, p, tid, FT);
                fft_pair_pass_fixed<FM, FM, 0u, true>(cx, t11, t9, FM);
                fft_pair_pass_fixed<FM, 128u, 4u, false>(cx, t7, t5, need);
                fft_tail_pass_fixed<FM, 8u, FT, false>(cx, t3, need);
            } else {
                /* three fused pairs, no stage left behind them */
                const double2 *t12 = p.tw_complex + p.tw_complex_off[12], *t10 = p.tw_complex + p.tw_complex_off[10], *t8 = p.tw_complex + p.tw_complex_off[8];
                const double2 *t6 = p.tw_complex + p.tw_complex_off[6], *t4 = p.tw_complex + p.tw_complex_off[4], *t2 = p.tw_complex + p.tw_complex_off[2];
                ws.first_pass_fixed(cx, t12, t10);
                fft_pair_pass_fixed<FM, 256u, 4u, true>(cx, t8, t6, FM);
                fft_pair_pass_fixed<FM, 16u, 8u, true>(cx, t4, t2, FM);
                real_split_power(cx, 2u * FM, p, tid, FT);
                fft_pair_pass_fixed<FM, FM, 0u, true>(cx, t12, t10, FM);
                fft_pair_pass_fixed<FM, 256u, 4u, true>(cx, t8, t6, FM);
                fft_pair_pass_fixed<FM, 16u, 8u, false>(cx, t4, t2, need);
            }
            for (uint32_t i = tid; i < nlags; i += FT) {
                const double2 e = cx[fft_slot(i >> 1)];
                lags[(size_t)i * lag_step] = (i < 2u * FM) ? ((i & 1u) ? -e.y : e.x) * ac_scale : 0.0;
            }
            __syncthreads();
            return;
        }
    }
    /* dir 0: forward transform of the windowed samples; dir 1: the inverse transform, evaluated as the
     * conjugate of a forward transform of conjugated data (see butterfly4) so both directions share one
     * copy of the pass code */
    #pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
        uint32_t nn = M, lgs = 0;
        const uint32_t need = dir ? ((want + 1u) >> 1) : M;
        if (dir == 0) {
            if (M >= 16u) {
                const uint32_t lgn = 31u - (uint32_t)__clz((int)nn);
                ws.first_pass(cx, M, nn, lgs, p.tw_complex + p.tw_complex_off[lgn], p.tw_complex + p.tw_complex_off[lgn - 2u], need);   /* windows the samples itself */
                nn >>= 4; lgs += 4;
            } else {
                for (uint32_t c = tid; c < M; c += nthreads) { cx[fft_slot(c)] = ws.element(c); }
                __syncthreads();
            }
        } else {
            real_split_power(cx, N, p, tid, nthreads);
        }
        #pragma unroll 1
        while (nn >= 16u) {
            const uint32_t lgn = 31u - (uint32_t)__clz((int)nn);
            fft_pair_pass<NoSource, false>(cx, M, nn, lgs, p.tw_complex + p.tw_complex_off[lgn], p.tw_complex + p.tw_complex_off[lgn - 2u], need, NoSource());
            nn >>= 4; lgs += 4;
        }
        fft_tail_pass(cx, M, nn, p.tw_complex + p.tw_complex_off[3], p.tw_complex + p.tw_complex_off[2], need);
    }
    /* the buffer holds the conjugate of the reference's inverse transform: odd lags are -imag */
    const double scale = ac_scale;
    if constexpr (kChain) {
        for (uint32_t e = tid; e < M; e += nthreads) { const double2 v = cx[fft_slot(e)]; pbuf[2u * e] = v.x; pbuf[2u * e + 1u] = -v.y; }
        __syncthreads();
        for (uint32_t i = tid; i < nlags; i += nthreads) { lags[(size_t)i * lag_step] = pbuf[i] * scale; }
    } else {
        for (uint32_t i = tid; i < nlags; i += nthreads) {
            double v = 0.0;
            if (i < N) { const double2 e = cx[fft_slot(i >> 1)]; v = ((i & 1u) ? -e.y : e.x) * scale; }
            lags[(size_t)i * lag_step] = v;
        }
    }
    __syncthreads();
}

/* correctly rounded s^-1/2 (the reference calls glibc pow(s, -0.5), lpc.c:591) */
__device__ __forceinline__ double inv_sqrt_cr(double s)
{
    const double y = 1.0 / sqrt(s);
    const double t = y * y, tl = __fma_rn(y, y, -t);
    const double q = s * t, ql = __fma_rn(s, t, -q);
    const double e = ((1.0 - q) - ql) - s * tl;          /* 1 - s*y^2 to ~2^-100 */
    return __fma_rn(0.5 * y, e, y);
}

/* geometric-distribution entropy (srla_encoder.c:873-885) */
__device__ __forceinline__ double geometric_entropy(double mean_abs, uint32_t bps)
{
    const double int_mean = mean_abs * (double)(1 << (bps - 1u));
    const double rho = 1.0 / (1.0 + int_mean);
    const double inv = 1.0 - rho;
    if (mean_abs < 1e-16) { return 0.0; }
    return -(inv * (log(inv) * 1.4426950408889634) + rho * (log(rho) * 1.4426950408889634)) / rho;
}

/* LTP pitch pick (lpc.c:1473-1555): serial, one thread */
__device__ int detect_pitch(const double *r, uint32_t *period)
{
    const uint32_t lo = kLtpMinPeriod, hi = kLtpMaxPeriod;
    uint32_t cand[20], ncand = 0, i = lo;
    double best_peak = 0.0;
    while (i < hi && ncand < 20) {
        uint32_t start, end, arg = 0; double peak = 0.0;
        for (start = i; start < hi; start++) { if (r[start - 1] < 0.0 && r[start] > 0.0) { break; } }
        for (end = start + 1; end < hi - 1; end++) { if (r[end] > 0.0 && r[end + 1] < 0.0) { break; } }
        for (uint32_t j = start; j <= end; j++) {
            if (r[j] > r[j - 1] && r[j] > r[j + 1] && r[j] > peak) { arg = j; peak = r[j]; }
        }
        if (arg) { cand[ncand++] = arg; if (peak > best_peak) { best_peak = peak; } }
        i = end + 1;
    }
    if (!ncand || best_peak < 0.1 * r[0]) { return 0; }
    for (uint32_t k = 0; k < ncand; k++) { if (r[cand[k]] >= 0.9 * best_peak) { *period = cand[k]; return 1; } }
    return 0;
}

/* The same pick by ONE WARP (front_kernel: 255 other threads wait for it).  The three comparisons every lag enters -- rising
 * zero crossing, falling zero crossing, local maximum -- are taken for all lags at once and kept as bit masks (`mask`: 27 words
 * of shared memory); lane 0 then walks the segments with find-first-set instead of two loops over the lags, and only looks at
 * the lags that ARE local maxima.  Same comparisons on the same values, same order of the candidates: the same period.
 * Returns (in every lane) 0 / 1 like detect_pitch; *period is set in every lane. */
__device__ int detect_pitch_warp(const double *r, uint32_t *period, uint32_t *mask, const uint32_t lane)
{
    const uint32_t lo = kLtpMinPeriod, hi = kLtpMaxPeriod;
    for (uint32_t w = 0; w < 9u; ++w) {
        const uint32_t j = 32u * w + lane;
        bool up = false, down = false, top = false;
        if (j >= 1u && j + 1u < (uint32_t)kLtpLags) {
            const double a = r[j - 1u], b = r[j], c = r[j + 1u];
            up = a < 0.0 && b > 0.0; down = b > 0.0 && c < 0.0; top = b > a && b > c;
        }
        const uint32_t mu = __ballot_sync(0xffffffffu, up), md = __ballot_sync(0xffffffffu, down), mt = __ballot_sync(0xffffffffu, top);
        if (lane == 0u) { mask[w] = mu; mask[9u + w] = md; mask[18u + w] = mt; }
    }
    __syncwarp();
    int found = 0; uint32_t result = 0;
    if (lane == 0u) {
        /* first set bit of a 288-bit mask at or behind `from`, or `none` */
        auto next_bit = [&](const uint32_t *m, uint32_t from, uint32_t none) {
            for (uint32_t w = from >> 5; w < 9u; ++w) {
                uint32_t bits = m[w];
                if (w == (from >> 5)) { bits &= 0xffffffffu << (from & 31u); }
                if (bits) { return 32u * w + (uint32_t)__ffs((int)bits) - 1u; }
            }
            return none;
        };
        uint32_t cand[20], ncand = 0, i = lo;
        double best_peak = 0.0;
        while (i < hi && ncand < 20u) {
            uint32_t start = next_bit(mask, i, hi);
            if (start > hi) { start = hi; }
            uint32_t end = (start + 1u < hi - 1u) ? next_bit(mask + 9, start + 1u, hi - 1u) : start + 1u;
            if (start + 1u < hi - 1u && end > hi - 1u) { end = hi - 1u; }
            uint32_t arg = 0; double peak = 0.0;
            for (uint32_t j = next_bit(mask + 18, start, 0xffffffffu); j <= end; j = next_bit(mask + 18, j + 1u, 0xffffffffu)) {
                if (r[j] > peak) { arg = j; peak = r[j]; }
            }
            if (arg) { cand[ncand++] = arg; if (peak > best_peak) { best_peak = peak; } }
            i = end + 1u;
        }
        if (ncand && !(best_peak < 0.1 * r[0])) {
            for (uint32_t k = 0; k < ncand; k++) { if (r[cand[k]] >= 0.9 * best_peak) { result = cand[k]; found = 1; break; } }
        }
    }
    found = __shfl_sync(0xffffffffu, found, 0); result = __shfl_sync(0xffffffffu, result, 0);
    *period = result;
    return found;
}

/* 3-tap (or 1-tap) normal equations by Cholesky (lpc.c:573-631, 1558-1649) + 6-bit quantisation
 * (srla_encoder.c:1032-1047).  returns 0 ok / 1 the reference would fail.  serial, one thread.
 * pitch: -1 = look for the pitch here (detect_pitch); 0 / 1 = the result of detect_pitch_warp, period in *period_out */
__device__ int ltp_solve(double *r, const uint32_t order, uint32_t *period_out, int32_t *qcoef, const int pitch = -1)
{
    double A[3][3], inv_diag[3], sol[3];
    uint32_t period = (pitch == 1) ? *period_out : 0u;
    const int dim = (int)order;
    *period_out = 0;
    if (fabs(r[0]) <= (double)FLT_MIN) { return 0; }                 /* lpc.c:1602 */
    if (pitch == 0 || (pitch < 0 && !detect_pitch(r, &period))) { return 0; }
    if (period < order / 2u + 1u) { return 0; }
    r[0] *= (1.0 + 1e-5);
    for (int i = 0; i < dim; i++) { for (int j = 0; j < dim; j++) { A[i][j] = r[(i > j) ? i - j : j - i]; } }
    for (int i = 0; i < dim; i++) {
        double s = A[i][i];
        for (int k = i - 1; k >= 0; k--) { s -= A[i][k] * A[i][k]; }
        if (s <= 0.0) { return 1; }
        inv_diag[i] = inv_sqrt_cr(s);
        for (int j = i + 1; j < dim; j++) {
            s = A[i][j];
            for (int k = i - 1; k >= 0; k--) { s -= A[i][k] * A[j][k]; }
            A[j][i] = s * inv_diag[i];
        }
    }
    const double *rhs = &r[period - order / 2u];
    for (int i = 0; i < dim; i++) {
        double s = rhs[i];
        for (int j = i - 1; j >= 0; j--) { s -= A[i][j] * sol[j]; }
        sol[i] = s * inv_diag[i];
    }
    for (int i = dim - 1; i >= 0; i--) {
        double s = sol[i];
        for (int j = i + 1; j < dim; j++) { s -= A[j][i] * sol[j]; }
        sol[i] = s * inv_diag[i];
    }
    for (int i = 0; i < dim; i++) {
        int32_t v = (int32_t)round_half_away(sol[i] * 32.0);
        if (v < -32) { v = -32; }
        if (v > 31) { v = 31; }
        qcoef[dim - 1 - i] = v;
    }
    *period_out = period;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Candidate signal preparation shared by front_kernel and residual_kernel
 * ---------------------------------------------------------------------------------------------- */
/* raw 16-byte (int32 PCM) or 8-byte (int16 PCM) global load of one sample quad, unpacked later: the loads of a
 * whole batch of quads are issued back to back so a thread waits for the memory latency once per batch */
__device__ __forceinline__ int4 load_quad_raw(const StreamDev &st, uint32_t ch, uint32_t idx)
{
    const unsigned long long at = (unsigned long long)ch * st.stride + idx;
    if (st.sample_bytes == 2u) { const int2 v = ldg_stream_v2(reinterpret_cast<const short *>(st.pcm) + at); return make_int4(v.x, v.y, 0, 0); }
    return ldg_stream_v4(reinterpret_cast<const int32_t *>(st.pcm) + at);
}
__device__ __forceinline__ int4 unpack_quad(const StreamDev &st, int4 v)
{
    if (st.sample_bytes == 2u) { return make_int4((int32_t)(short)(v.x & 0xffff), v.x >> 16, (int32_t)(short)(v.y & 0xffff), v.y >> 16); }
    return v;
}

/* load, >> offset_lshift, mid/side (srla_encoder.c:1229-1253, srla_utility.c:91-103) -> raw[0..n).
 * returns OR of the unshifted samples of a plain channel candidate (0 for M/S). */
template <int kBatch, bool kShift>            /* quads in flight per thread; kShift: the stream has a non-zero offset shift */
__device__ __forceinline__ int load_candidate_impl(const StreamDev &st, const Job &job, const LaunchParams &p, uint32_t cand,
                                                   uint32_t lshift_in, int32_t *raw)
{
    const uint32_t lshift = kShift ? lshift_in : 0u;          /* a compile-time zero removes every shift below */
    const uint32_t n = job.nsmpl;
    const uint32_t first_ch = (p.nch >= 2u) ? 2u : 0u;
    const bool ms = (p.nch >= 2u) && (cand < 2u);
    const uint32_t ch = ms ? 0u : cand - first_ch;
    const bool vec = quad_aligned(st, ch, job.offset) && (!ms || quad_aligned(st, 1u, job.offset));
    const uint32_t nquad = vec ? (n >> 2) : 0u;
    const uint32_t T = blockDim.x;
    int nz = 0;
    if (ms) {
        for (uint32_t g0 = threadIdx.x; g0 < nquad; g0 += kBatch * T) {
            int4 lraw[kBatch], rraw[kBatch];
            #pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t g = g0 + (uint32_t)b * T;
                if (g < nquad) { lraw[b] = load_quad_raw(st, 0, job.offset + 4u * g); rraw[b] = load_quad_raw(st, 1, job.offset + 4u * g); }
            }
            #pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t g = g0 + (uint32_t)b * T;
                if (g < nquad) {
                    const int4 lq = unpack_quad(st, lraw[b]), rq = unpack_quad(st, rraw[b]);
                    const int32_t l[4] = { asr32(lq.x, lshift), asr32(lq.y, lshift), asr32(lq.z, lshift), asr32(lq.w, lshift) };
                    const int32_t r[4] = { asr32(rq.x, lshift), asr32(rq.y, lshift), asr32(rq.z, lshift), asr32(rq.w, lshift) };
                    int32_t o[4];
                    #pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int32_t side = (int32_t)((uint32_t)r[t] - (uint32_t)l[t]);
                        o[t] = (cand == 1u) ? side : (int32_t)((uint32_t)l[t] + (uint32_t)(side >> 1));
                    }
                    *reinterpret_cast<int4 *>(raw + 4u * g) = make_int4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        for (uint32_t i = 4u * nquad + threadIdx.x; i < n; i += T) {
            const int32_t l = asr32(load_sample(st, 0, job.offset + i), lshift);
            const int32_t r = asr32(load_sample(st, 1, job.offset + i), lshift);
            const int32_t side = (int32_t)((uint32_t)r - (uint32_t)l);
            raw[i] = (cand == 1u) ? side : (int32_t)((uint32_t)l + (uint32_t)(side >> 1));
        }
    } else {
        for (uint32_t g0 = threadIdx.x; g0 < nquad; g0 += kBatch * T) {
            int4 qraw[kBatch];
            #pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t g = g0 + (uint32_t)b * T;
                if (g < nquad) { qraw[b] = load_quad_raw(st, ch, job.offset + 4u * g); }
            }
            #pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t g = g0 + (uint32_t)b * T;
                if (g < nquad) {
                    const int4 q = unpack_quad(st, qraw[b]);
                    nz |= q.x | q.y | q.z | q.w;
                    *reinterpret_cast<int4 *>(raw + 4u * g) = make_int4(asr32(q.x, lshift), asr32(q.y, lshift), asr32(q.z, lshift), asr32(q.w, lshift));
                }
            }
        }
        for (uint32_t i = 4u * nquad + threadIdx.x; i < n; i += T) {
            const int32_t v = load_sample(st, ch, job.offset + i);
            nz |= v;
            raw[i] = asr32(v, lshift);
        }
    }
    return nz;
}

/* almost every stream has offset shift 0 (its first odd sample comes early): that case runs without the variable shifts */
template <int kBatch>
__device__ __forceinline__ int load_candidate(const StreamDev &st, const Job &job, const LaunchParams &p, uint32_t cand,
                                              uint32_t lshift, int32_t *raw)
{
    return (lshift == 0u) ? load_candidate_impl<kBatch, false>(st, job, p, cand, 0u, raw) : load_candidate_impl<kBatch, true>(st, job, p, cand, lshift, raw);
}

/* pre-emphasis (srla_utility.c:342-358), filter memory seeded with the first sample; raw -> sig,
 * plus the zero padding the FIR's vector loads may touch */
__device__ __forceinline__ void apply_preemphasis(const int32_t *raw, int32_t *sig, uint32_t n, int32_t pre_coef)
{
    const uint32_t c = (uint32_t)pre_coef, nquad = n >> 2;
    for (uint32_t g = threadIdx.x; g < nquad; g += blockDim.x) {
        const int4 q = *reinterpret_cast<const int4 *>(raw + 4u * g);
        const int32_t prv = raw[g ? 4u * g - 1u : 0u];
        int4 o;
        o.x = (int32_t)((uint32_t)q.x - (uint32_t)((int32_t)((uint32_t)prv * c) >> 4));
        o.y = (int32_t)((uint32_t)q.y - (uint32_t)((int32_t)((uint32_t)q.x * c) >> 4));
        o.z = (int32_t)((uint32_t)q.z - (uint32_t)((int32_t)((uint32_t)q.y * c) >> 4));
        o.w = (int32_t)((uint32_t)q.w - (uint32_t)((int32_t)((uint32_t)q.z * c) >> 4));
        *reinterpret_cast<int4 *>(sig + 4u * g) = o;
    }
    for (uint32_t i = 4u * nquad + threadIdx.x; i < n; i += blockDim.x) {
        const int32_t cur = raw[i], prv = raw[(i == 0u) ? 0u : i - 1u];
        sig[i] = (int32_t)((uint32_t)cur - (uint32_t)((int32_t)((uint32_t)prv * c) >> 4));
    }
    if (threadIdx.x < 4) { sig[-1 - (int)threadIdx.x] = 0; }
    for (uint32_t i = n + threadIdx.x; i < round_up_u32(n, 4) + 12u; i += blockDim.x) { sig[i] = 0; }
}

/* long-term prediction residual replaces the signal (srla_lpc_predict.c:267-294); tmp: n int32 */
__device__ __forceinline__ void apply_ltp(int32_t *sig, int32_t *tmp, uint32_t n, uint32_t ltp_order, uint32_t period,
                                          int32_t c0, int32_t c1, int32_t c2)
{
    const uint32_t half_order = ltp_order >> 1;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        int32_t v = sig[i];
        if (i >= period + half_order + 1u) {
            const int32_t *x = sig + (i - period - half_order);
            uint32_t acc = 16u;
            acc += (uint32_t)c0 * (uint32_t)x[0];
            if (ltp_order > 1u) { acc += (uint32_t)c1 * (uint32_t)x[1]; acc += (uint32_t)c2 * (uint32_t)x[2]; }
            v = (int32_t)((uint32_t)v - (uint32_t)((int32_t)acc >> 5));
        }
        tmp[i] = v;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { sig[i] = tmp[i]; }
    __syncthreads();
}

/* ------------------------------------------------------------------------------------------------
 * front_kernel: one CTA per (job, candidate).  Pre-emphasis decision, optional LTP analysis, and
 * the Welch-windowed FFT autocorrelation of the signal the LPC stage sees; lags 0..P go to HBM.
 * ---------------------------------------------------------------------------------------------- */
/* 16-bit pair entry of the filtered samples x[0..4] = signal[4j .. 4j+4] (see residual_kernel) */
__device__ __forceinline__ int4 pack_pair_entry(const int32_t x[5])
{
    int4 z;
    z.x = (int32_t)(((uint32_t)x[0] & 0xffffu) | ((uint32_t)x[1] << 16));
    z.y = (int32_t)(((uint32_t)x[2] & 0xffffu) | ((uint32_t)x[3] << 16));
    z.z = (int32_t)(((uint32_t)x[1] & 0xffffu) | ((uint32_t)x[2] << 16));
    z.w = (int32_t)(((uint32_t)x[3] & 0xffffu) | ((uint32_t)x[4] << 16));
    return z;
}

/* pre-emphasis coefficient (srla_utility.c:214-257) from the candidate samples raw[0..n): r0, r1 as exact
 * integers.  Every thread returns the coefficient; thread 0 also records it. */
template <int kT>
__device__ __forceinline__ int32_t preemphasis_coefficient(const int32_t *raw, const uint32_t n, CandOut *out,
                                                           unsigned long long *red64, int32_t *sh_coef)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long r0 = 0, r1 = 0;
    const uint32_t nquad = n >> 2;
    for (uint32_t g = tid; g < nquad; g += kT) {
        const int4 q = *reinterpret_cast<const int4 *>(raw + 4u * g);
        const long long a = q.x, b = q.y, c = q.z, d = q.w;
        r0 += a * a + b * b + c * c + d * d;
        r1 += a * b + b * c + c * d;
        if (4u * g + 4u < n) { r1 += d * (long long)raw[4u * g + 4u]; }
    }
    for (uint32_t i = 4u * nquad + tid; i < n; i += kT) {
        const long long a = raw[i];
        r0 += a * a;
        if (i + 1u < n) { r1 += a * (long long)raw[i + 1u]; }
    }
    r0 = warp_sum_ll(r0); r1 = warp_sum_ll(r1);
    if (lane == 0) { red64[2 * warp] = (unsigned long long)r0; red64[2 * warp + 1] = (unsigned long long)r1; }
    __syncthreads();
    if (tid == 0) {
        long long s0 = 0, s1 = 0;
        for (int w = 0; w < kT / 32; ++w) { s0 += (long long)red64[2 * w]; s1 += (long long)red64[2 * w + 1]; }
        int32_t c = 0;
        if (s0 != 0) {
            double v = ((double)s1 / (double)s0) * 16.0;
            if (s0 >= (1ll << 53)) {
                /* the reference's sequential double sums round above 2^53 (relative error
                 * <= n * 2^-53 each); only a value this close to a rounding boundary can differ */
                const double frac = fabs(v) - floor(fabs(v));
                if (fabs(frac - 0.5) < 1e-7) {
                    double q0 = 0.0, q1 = 0.0;
                    for (uint32_t i = 0; i + 1u < n; ++i) { const double a = raw[i], b = raw[i + 1u]; q0 += a * a; q1 += a * b; }
                    { const double a = raw[n - 1u]; q0 += a * a; }
                    v = (q1 / q0) * 16.0;
                }
            }
            c = (int32_t)round_half_away(v);
            if (c < -16) { c = -16; }
            if (c > 15) { c = 15; }
        }
        *sh_coef = c;
        out->pre_coef = c; out->pre_prev = raw[0];
    }
    __syncthreads();
    return *sh_coef;
}

/* kOcc: resident CTAs per SM the register allocation is sized for */
template <int kT, int kOcc, bool kLtp>
__global__ void __launch_bounds__(kT, kOcc) front_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const FrontLayout L = make_front_layout(p.nmax, p.fft_max, p.ltp_order);
    double   *region_d = reinterpret_cast<double *>(smem + L.region_off);
    int32_t  *region_i = reinterpret_cast<int32_t *>(smem + L.region_off);
    int32_t  *sig      = reinterpret_cast<int32_t *>(smem + L.sig_off) + 4;     /* 4 ints of front padding */
    double   *lags     = reinterpret_cast<double *>(smem + L.lags_off);
    __shared__ unsigned long long red64[2 * (kT / 32)];
    __shared__ int32_t  sh_i[8];
    __shared__ uint32_t sh_u[8];

    const int tid = threadIdx.x;
    const uint32_t job_id = blockIdx.x / p.ncand, cand = blockIdx.x % p.ncand;
    const Job &job = p.jobs[job_id];
    const StreamDev &st = p.streams[job.stream];
    const uint32_t n = job.nsmpl, P = p.max_order;
    const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;
    CandOut *out = p.cand + (size_t)job_id * p.ncand + cand;

    /* without LTP the candidate goes straight into the signal buffer and the pre-emphasis is folded into
     * the window pass; with LTP the filtered signal itself is needed in shared memory */
    int32_t *raw = kLtp ? region_i : sig;
    int nz = load_candidate<8>(st, job, p, cand, lshift, raw);          /* 4096 samples / 128 threads = 8 quads each */
    nz = __syncthreads_or(nz);
    if (tid == 0) {
        out->nonzero = (nz != 0); out->status = 0; out->order = 0; out->rshift = 0;
        out->ltp_period = 0; out->ltp_coef[0] = 0; out->ltp_coef[1] = 0; out->ltp_coef[2] = 0;
        out->total_bits = 0; out->residual_bits = 0; out->pre_coef = 0; out->pre_prev = 0;
    }
    if (n <= P) { return; }                          /* RAW block (srla_encoder.c:777-779): nothing to analyse */

    const int32_t pre_coef = preemphasis_coefficient<kT>(raw, n, out, red64, &sh_i[0]);
    /* lags of 32 consecutive candidates are interleaved for the lpc kernel: [lag][candidate % 32] */
    double *g = p.lags + (size_t)(blockIdx.x >> 5) * p.lag_stride * 32u + (blockIdx.x & 31u);
    if (!kLtp) {
        /* ---- autocorrelation of the signal the LPC stage sees (lpc.c:444-483) ---- */
        if (P > 0u) { welch_autocorr<true>(raw, pre_coef, n, region_d, g, 32u, P + 1u, job, p); }
        return;
    }
    apply_preemphasis(region_i, sig, n, pre_coef);
    __syncthreads();

    /* ---- long-term prediction (srla_encoder.c:1008-1058) ---- */
    welch_autocorr<false>(sig, 0, n, region_d, lags, 1u, kLtpLags, job, p);
    /* lags 0..262 come from the transform; 263.. are never written by the reference (zero pages) */
    if ((uint32_t)tid < (uint32_t)kLtpLags - (kLtpMaxPeriod + 1u)) { lags[kLtpMaxPeriod + 1u + (uint32_t)tid] = 0.0; }
    __syncthreads();
    __shared__ uint32_t pitch_mask[27];
    uint32_t period = 0; int pitch = 0;
    if (tid < 32) { pitch = detect_pitch_warp(lags, &period, pitch_mask, (uint32_t)tid); }
    if (tid == 0) {
        int32_t q[3] = { 0, 0, 0 };
        const int rc = ltp_solve(lags, p.ltp_order, &period, q, pitch);
        sh_u[0] = period; sh_u[1] = (uint32_t)rc; sh_i[1] = q[0]; sh_i[2] = q[1]; sh_i[3] = q[2];
        out->status = (uint32_t)rc;
        if (!rc && period > 0u) { out->ltp_period = period; out->ltp_coef[0] = q[0]; out->ltp_coef[1] = q[1]; out->ltp_coef[2] = q[2]; }
    }
    __syncthreads();
    if (sh_u[1]) { return; }
    if (sh_u[0] > 0u) { apply_ltp(sig, region_i, n, p.ltp_order, sh_u[0], sh_i[1], sh_i[2], sh_i[3]); }
    if (P > 0u) { welch_autocorr<false>(sig, 0, n, region_d, g, 32u, P + 1u, job, p); }
}

/* ------------------------------------------------------------------------------------------------
 * front_tail_kernel: the reference's stale-scratch corners (lpc.c:260-264, 371-373), fixed blocks only.
 * The reference's LPC calculator keeps ONE scratch buffer for all its calls.  For an odd block length the Welch window
 * leaves the middle sample of that buffer as the previous call's inverse transform left it, and with LTP on a block whose
 * transform is shorter than 263 points the lags beyond the transform are copied from it as well.  With fixed blocks only
 * the last block of a stream can be that short or odd (the block size itself must be even and, with LTP, at least 263),
 * so the chain of calls that matters is short: the last call of the latest analysed block in front of it (silent blocks
 * make no call), then the calls of the tail block's own candidates in the reference's order M, S, channel 0, 1, ...
 * (srla_encoder.c:1254-1273; two calls per candidate with LTP).  One CTA replays that chain for one tail job with `pbuf`
 * standing in for the scratch buffer (zeroed first: a handle on fresh memory, like the `srla` CLI's) and overwrites what
 * front_kernel wrote for the job's candidates.
 * ---------------------------------------------------------------------------------------------- */
/* job / first: positions in the CALL LIST the kernel is given (the reference's analysis calls of the stream in order; a
 * call's predecessors are searched from job - 1 down to first); out: the job's index in the launch's own job list */
struct TailJob { uint32_t job; uint32_t first_of_stream; uint32_t out; };

template <int kT, bool kLtp>
__device__ void front_chain_step(const LaunchParams &p, const Job &job, const StreamDev &st, const uint32_t cand, const uint32_t lshift,
                                 const FrontLayout &L, unsigned char *smem, double *pbuf, CandOut *out, double *g, const uint32_t gstep,
                                 unsigned long long *red64, int32_t *sh_i, uint32_t *sh_u)
{
    double   *region_d = reinterpret_cast<double *>(smem + L.region_off);
    int32_t  *region_i = reinterpret_cast<int32_t *>(smem + L.region_off);
    int32_t  *sig      = reinterpret_cast<int32_t *>(smem + L.sig_off) + 4;
    double   *lags     = reinterpret_cast<double *>(smem + L.lags_off);
    const int tid = threadIdx.x;
    const uint32_t n = job.nsmpl, P = p.max_order;
    __syncthreads();
    /* zero padding of the signal buffer the window pass may touch (a longer block was here before) */
    for (uint32_t i = n + tid; i < round_up_u32(p.nmax, 4) + 12u; i += kT) { sig[i] = 0; }
    int32_t *raw = kLtp ? region_i : sig;
    int nz = load_candidate<8>(st, job, p, cand, lshift, raw);
    nz = __syncthreads_or(nz);
    if (tid == 0) {
        out->nonzero = (nz != 0); out->status = 0; out->order = 0; out->rshift = 0;
        out->ltp_period = 0; out->ltp_coef[0] = 0; out->ltp_coef[1] = 0; out->ltp_coef[2] = 0;
        out->total_bits = 0; out->residual_bits = 0; out->pre_coef = 0; out->pre_prev = 0;
    }
    const int32_t pre_coef = preemphasis_coefficient<kT>(raw, n, out, red64, &sh_i[0]);
    if (!kLtp) {
        if (P > 0u) { welch_autocorr<true, true>(raw, pre_coef, n, region_d, g, gstep, P + 1u, job, p, pbuf); }
        return;
    }
    apply_preemphasis(region_i, sig, n, pre_coef);
    __syncthreads();
    welch_autocorr<false, true>(sig, 0, n, region_d, lags, 1u, kLtpMaxPeriod + 1u, job, p, pbuf);
    if (tid == 0) {
        for (uint32_t i = kLtpMaxPeriod + 1u; i < (uint32_t)kLtpLags; ++i) { lags[i] = 0.0; }      /* never written by any call */
        uint32_t period = 0; int32_t q[3] = { 0, 0, 0 };
        const int rc = ltp_solve(lags, p.ltp_order, &period, q);
        sh_u[0] = period; sh_u[1] = (uint32_t)rc; sh_i[1] = q[0]; sh_i[2] = q[1]; sh_i[3] = q[2];
        out->status = (uint32_t)rc;
        if (!rc && period > 0u) { out->ltp_period = period; out->ltp_coef[0] = q[0]; out->ltp_coef[1] = q[1]; out->ltp_coef[2] = q[2]; }
    }
    __syncthreads();
    if (sh_u[1]) { return; }
    if (sh_u[0] > 0u) { apply_ltp(sig, region_i, n, p.ltp_order, sh_u[0], sh_i[1], sh_i[2], sh_i[3]); }
    if (P > 0u) { welch_autocorr<false, true>(sig, 0, n, region_d, g, gstep, P + 1u, job, p, pbuf); }
}

template <int kT, bool kLtp>
__global__ void __launch_bounds__(kT) front_tail_kernel(const __grid_constant__ LaunchParams p, const TailJob *tails, const Job *jobs_all,
                                                        const uint32_t group_first, const uint32_t pbuf_len, double *pbuf_global)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const FrontLayout L = make_front_layout(p.nmax, p.fft_max, p.ltp_order);
    /* the scratch-buffer replica lives behind the kernel's own shared memory, or -- for block sizes whose transform already
     * fills it -- in global memory (pbuf_len + 272 doubles per CTA) */
    double *pbuf = pbuf_global ? pbuf_global + (size_t)blockIdx.x * (pbuf_len + 272u) : reinterpret_cast<double *>(smem + L.total);
    __shared__ unsigned long long red64[2 * (kT / 32)];
    __shared__ int32_t  sh_i[8];
    __shared__ uint32_t sh_u[8];
    __shared__ CandOut scratch_out;
    const int tid = threadIdx.x;
    const TailJob tj = tails[blockIdx.x];
    const Job job = jobs_all[tj.job];
    const StreamDev st = p.streams[job.stream];
    const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;
    const uint32_t P = p.max_order;
    if (job.nsmpl <= P) { return; }
    for (uint32_t i = tid; i < pbuf_len; i += kT) { pbuf[i] = 0.0; }
    /* A call depends on what earlier calls left in the scratch buffer when its length is odd (the middle sample of the window)
     * or, with LTP, when its transform is shorter than the 263 lags the pitch search reads.  What it finds there was written
     * by the LATEST earlier call whose transform covers the index: a block with more samples than the order that is not
     * silent (srla_encoder.c:766-796).  If that call depends on history itself (variable-block search: the clipped segments
     * at the end of a stream follow one another), its own predecessor is looked up the same way: `chain` collects those
     * calls, newest first, until one is reached that does not depend on anything (`seed`). */
    auto depends = [&](const Job &j) { return (j.nsmpl & 1u) != 0u || (kLtp && ceil_pow2_u32(j.nsmpl) < (uint32_t)kLtpMaxPeriod + 1u); };
    auto reach = [&](const Job &j) {                    /* the highest scratch index the call reads before writing it */
        uint32_t r = (j.nsmpl & 1u) ? ((j.nsmpl - 1u) >> 1) : 0u;
        if (kLtp && ceil_pow2_u32(j.nsmpl) < (uint32_t)kLtpMaxPeriod + 1u) { r = max(r, ceil_pow2_u32(j.nsmpl)); }
        return r;
    };
    constexpr int kMaxChain = 12;
    uint32_t chain[kMaxChain]; int depth = 0;
    uint32_t seed = 0xffffffffu;
    {
        uint32_t pos = tj.job;
        Job cur = job;
        for (;;) {
            const uint32_t need_idx = reach(cur);
            uint32_t pred = pos;
            bool found = false;
            while (!found && pred > tj.first_of_stream) {
                pred--;
                const Job pj = jobs_all[pred];
                int nz = 0;
                if (pj.nsmpl > P && ceil_pow2_u32(pj.nsmpl) > need_idx) {
                    for (uint32_t ch = 0; ch < p.nch; ++ch) { for (uint32_t i = tid; i < pj.nsmpl; i += kT) { nz |= load_sample(st, ch, pj.offset + i); } }
                }
                found = __syncthreads_or(nz) != 0;
            }
            if (!found) { break; }
            const Job pj = jobs_all[pred];
            if (!depends(pj) || depth == kMaxChain) { seed = pred; break; }
            chain[depth++] = pred; pos = pred; cur = pj;
        }
    }
    if (seed != 0xffffffffu) {
        /* its last call: the last candidate (the reference analyses M, S first, then the channels in order) */
        const Job pj = jobs_all[seed];
        front_chain_step<kT, kLtp>(p, pj, st, p.ncand - 1u, lshift, L, smem, pbuf, &scratch_out, pbuf + pbuf_len, 1u, red64, sh_i, sh_u);   /* lags to a dump area */
    }
    for (int d = depth - 1; d >= 0; --d) {
        const Job cj = jobs_all[chain[d]];
        for (uint32_t cand = 0; cand < p.ncand; ++cand) {
            front_chain_step<kT, kLtp>(p, cj, st, cand, lshift, L, smem, pbuf, &scratch_out, pbuf + pbuf_len, 1u, red64, sh_i, sh_u);
        }
    }
    const uint32_t rel = tj.out - group_first;
    for (uint32_t cand = 0; cand < p.ncand; ++cand) {
        const uint32_t idx = rel * p.ncand + cand;
        double *g = p.lags + (size_t)(idx >> 5) * p.lag_stride * 32u + (idx & 31u);
        front_chain_step<kT, kLtp>(p, job, st, cand, lshift, L, smem, pbuf, p.cand + idx, g, 32u, red64, sh_i, sh_u);
    }
}

/* ------------------------------------------------------------------------------------------------
 * front_big_kernel: blocks of more than kMaxSharedBlock samples (the format's block header carries up to 65535,
 * srla_encoder.c:1593; the reference CLI admits -B < 65536, srla_codec.c:354).  Their transform does not fit an SM's
 * shared memory, so one CTA works on a job in global memory (the working set of the resident CTAs stays in L2): the
 * reference's own loop structure -- radix-4 Stockham stages ping-ponging between two buffers, one radix-2 stage at the
 * end (fft.c:71-128), the real-transform split (fft.c:147-198) -- with the butterflies of a stage spread over the threads.
 * Every butterfly is the expression tree front_kernel evaluates (butterfly4, the host-tabulated twiddle recurrence), the
 * inverse transform again runs as the conjugate of a forward transform of conjugated data, so the lags are bit-identical.
 * A CTA takes a whole job and analyses its candidates in the reference's order with `pbuf` standing in for the LPC
 * calculator's scratch buffer, exactly like front_tail_kernel; the stale middle sample / stale lags are only USED for the
 * jobs front_tail_kernel would replay (chain == true), so both paths deviate from the reference in the same documented
 * places and nowhere else.
 * ---------------------------------------------------------------------------------------------- */
/* forward complex FFT of M points, natural order, out of place between x and y; returns the buffer holding the result */
__device__ double2 *fft_stockham_global(double2 *x, double2 *y, const uint32_t M, const LaunchParams &p)
{
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    uint32_t n = M, lgs = 0;
    while (n > 2u) {
        const uint32_t n1 = n >> 2, lgn = 31u - (uint32_t)__clz((int)n);
        const double2 *tw = p.tw_complex + p.tw_complex_off[lgn];
        const uint32_t s = 1u << lgs, count = n1 << lgs;
        for (uint32_t b = tid; b < count; b += T) {
            const uint32_t pp = b >> lgs, q = b & (s - 1u);
            const Twiddle3 w = load_twiddle(tw, n1, pp);
            double2 y0, y1, y2, y3;
            butterfly4(x[q + ((pp) << lgs)], x[q + ((pp + n1) << lgs)], x[q + ((pp + 2u * n1) << lgs)], x[q + ((pp + 3u * n1) << lgs)], w, y0, y1, y2, y3);
            double2 *o = y + q + ((4u * pp) << lgs);
            o[0] = y0; o[s] = y1; o[2u * s] = y2; o[3u * s] = y3;
        }
        __syncthreads();
        n >>= 2; lgs += 2;
        double2 *t = x; x = y; y = t;
    }
    if (n == 2u) {
        const uint32_t s = 1u << lgs;
        for (uint32_t q = tid; q < s; q += T) { const double2 a = x[q], b = x[q + s]; y[q] = cadd(a, b); y[q + s] = csub(a, b); }
        __syncthreads();
        double2 *t = x; x = y; y = t;
    }
    return x;
}

/* Welch window + FFT autocorrelation of sig[0..n) (global memory), lags[0..nlags) out; kPre as in welch_autocorr.
 * pbuf: natural-order replica of the reference's scratch buffer (always updated); `chain`: its stale contents are used
 * (odd middle sample, lags beyond the transform) */
template <bool kPre>
__device__ void welch_autocorr_big(const int32_t *sig, const int32_t pre_coef, const uint32_t n, double2 *bufa, double2 *bufb, double *lags, const uint32_t lag_step,
                                   const uint32_t nlags, const Job &job, const LaunchParams &p, double *pbuf, const bool chain)
{
    const uint32_t tid = threadIdx.x, T = blockDim.x;
    const uint32_t N = ceil_pow2_u32(n), M = N >> 1;
    WindowSource<kPre, true> ws;
    ws.sig = sig; ws.n = n; ws.half_n = n >> 1; ws.pc = (uint32_t)pre_coef; ws.unit = 1.0; ws.div = job.welch_div * p.unit; ws.dn1 = (double)(int32_t)(n - 1u);
    ws.full = false;
    /* without the chain an odd block's middle sample is windowed like every other one (front_kernel's behaviour) */
    {
        const uint32_t mid = n >> 1;
        double stale = pbuf[mid];
        if (!chain && (n & 1u)) {
            const int32_t cur = sig[mid], prv = sig[mid ? mid - 1u : 0u];
            const int32_t x = kPre ? (int32_t)((uint32_t)cur - (uint32_t)((int32_t)((uint32_t)prv * ws.pc) >> 4)) : cur;
            const double ds = int_to_double((int32_t)mid);
            stale = int_to_double(x) * (ws.div * ds * (ws.dn1 - ds));
        }
        ws.stale = stale;
    }
    __syncthreads();
    for (uint32_t e = tid; e < M; e += T) { bufa[e] = ws.element(e); }
    __syncthreads();
    double2 *cx = fft_stockham_global(bufa, bufb, M, p);
    double2 *other = (cx == bufa) ? bufb : bufa;
    /* forward split, power spectrum, inverse split (stored conjugated): see welch_autocorr */
    {
        const uint32_t lgN = 31u - (uint32_t)__clz((int)N);
        const double2 *tw = p.tw_real + p.tw_real_off[lgN];
        const uint32_t quarter = N >> 2;
        for (uint32_t i = 1u + tid; i <= quarter; i += T) {
            const double2 w = __ldg(tw + (i - 1u));
            const double wr = w.x, wi_f = w.y, wi_b = -w.y;
            const uint32_t lo = i, hi = M - i;
            const double2 xl = cx[lo], xh = cx[hi];
            const double c2 = -0.5;
            const double h1r = 0.5 * (xl.x + xh.x), h1i = 0.5 * (xl.y - xh.y), h2r = -c2 * (xl.y + xh.y), h2i = c2 * (xl.x - xh.x);
            const double f1 = h1r + (wr * h2r) - (wi_f * h2i), f2 = h1i + (wr * h2i) + (wi_f * h2r);
            const double f3 = h1r - (wr * h2r) + (wi_f * h2i), f4 = -h1i + (wr * h2i) + (wi_f * h2r);
            double p_lo, p_hi;
            if (lo == hi) { p_hi = f3 * f3 + f4 * f4; p_lo = p_hi; } else { p_lo = f1 * f1 + f2 * f2; p_hi = f3 * f3 + f4 * f4; }
            const double g1r = 0.5 * (p_lo + p_hi), g2i = 0.5 * (p_lo - p_hi);
            const double t = wi_b * g2i, g2 = wr * g2i;
            if (lo != hi) { cx[lo] = make_double2(g1r - t, -g2); }
            cx[hi] = make_double2(g1r + t, -g2);
        }
        if (tid == 0) {
            const double2 dc = cx[0];
            const double f0 = dc.x + dc.y, f1 = dc.x - dc.y;
            const double q0 = f0 * f0, q1 = f1 * f1;
            cx[0] = make_double2(0.5 * (q0 + q1), -(0.5 * (q0 - q1)));
        }
        __syncthreads();
    }
    double2 *res = fft_stockham_global(cx, other, M, p);
    for (uint32_t e = tid; e < M; e += T) { const double2 v = res[e]; pbuf[2u * e] = v.x; pbuf[2u * e + 1u] = -v.y; }
    __syncthreads();
    const double scale = job.ac_scale;
    for (uint32_t i = tid; i < nlags; i += T) { lags[(size_t)i * lag_step] = (chain || i < N) ? pbuf[i] * scale : 0.0; }
    __syncthreads();
}

template <bool kLtp>
__global__ void __launch_bounds__(1024) front_big_kernel(const __grid_constant__ LaunchParams p)
{
    constexpr int kT = 1024;
    const FrontBigLayout L = make_front_big_layout(p.nmax, p.fft_max);
    unsigned char *base = p.big_scratch + (size_t)blockIdx.x * p.big_stride;
    int32_t *raw = reinterpret_cast<int32_t *>(base + L.raw_off) + 8;
    int32_t *sig = reinterpret_cast<int32_t *>(base + L.sig_off) + 8;
    double2 *bufa = reinterpret_cast<double2 *>(base + L.a_off), *bufb = reinterpret_cast<double2 *>(base + L.b_off);
    double *pbuf = reinterpret_cast<double *>(base + L.pbuf_off);
    double *lags = reinterpret_cast<double *>(base + L.lags_off);
    __shared__ unsigned long long red64[2 * (kT / 32)];
    __shared__ int32_t  sh_i[8];
    __shared__ uint32_t sh_u[8];
    __shared__ CandOut scratch_out;
    const int tid = threadIdx.x;
    const uint32_t P = p.max_order, pbuf_len = p.fft_max + 272u;

    /* one candidate of one job, chain semantics as in front_chain_step */
    auto step = [&](const Job &job, const StreamDev &st, uint32_t cand, uint32_t lshift, CandOut *out, double *g, uint32_t gstep, bool chain) {
        const uint32_t n = job.nsmpl;
        __syncthreads();
        int nz = load_candidate<4>(st, job, p, cand, lshift, raw);
        nz = __syncthreads_or(nz);
        if (tid == 0) {
            out->nonzero = (nz != 0); out->status = 0; out->order = 0; out->rshift = 0;
            out->ltp_period = 0; out->ltp_coef[0] = 0; out->ltp_coef[1] = 0; out->ltp_coef[2] = 0;
            out->total_bits = 0; out->residual_bits = 0; out->pre_coef = 0; out->pre_prev = 0;
        }
        const int32_t pre_coef = preemphasis_coefficient<kT>(raw, n, out, red64, &sh_i[0]);
        if (!kLtp) {
            if (P > 0u) { welch_autocorr_big<true>(raw, pre_coef, n, bufa, bufb, g, gstep, P + 1u, job, p, pbuf, chain); }
            return;
        }
        apply_preemphasis(raw, sig, n, pre_coef);
        __syncthreads();
        welch_autocorr_big<false>(sig, 0, n, bufa, bufb, lags, 1u, kLtpMaxPeriod + 1u, job, p, pbuf, chain);
        if (tid == 0) {
            for (uint32_t i = kLtpMaxPeriod + 1u; i < (uint32_t)kLtpLags; ++i) { lags[i] = 0.0; }
            uint32_t period = 0; int32_t q[3] = { 0, 0, 0 };
            const int rc = ltp_solve(lags, p.ltp_order, &period, q);
            sh_u[0] = period; sh_u[1] = (uint32_t)rc; sh_i[1] = q[0]; sh_i[2] = q[1]; sh_i[3] = q[2];
            out->status = (uint32_t)rc;
            if (!rc && period > 0u) { out->ltp_period = period; out->ltp_coef[0] = q[0]; out->ltp_coef[1] = q[1]; out->ltp_coef[2] = q[2]; }
        }
        __syncthreads();
        if (sh_u[1]) { return; }
        if (sh_u[0] > 0u) { apply_ltp(sig, raw, n, p.ltp_order, sh_u[0], sh_i[1], sh_i[2], sh_i[3]); }
        if (P > 0u) { welch_autocorr_big<false>(sig, 0, n, bufa, bufb, g, gstep, P + 1u, job, p, pbuf, chain); }
    };

    /* candidates of a block the reference makes no call for (RAW: not more samples than the order; SILENT): only the
     * non-zero flags decide_kernel reads */
    auto flags_only = [&](const Job &job, const StreamDev &st, uint32_t lshift, uint32_t j) {
        for (uint32_t cand = 0; cand < p.ncand; ++cand) {
            int nz = load_candidate<4>(st, job, p, cand, lshift, raw);
            nz = __syncthreads_or(nz);
            if (tid == 0) { CandOut *o = p.cand + (size_t)j * p.ncand + cand; o->nonzero = (nz != 0); o->status = 0; o->order = 0; o->rshift = 0; o->ltp_period = 0; o->total_bits = 0; o->residual_bits = 0; o->pre_coef = 0; o->pre_prev = 0; }
            __syncthreads();
        }
    };
    auto block_is_silent = [&](const Job &job, const StreamDev &st) {
        int nz = 0;
        for (uint32_t ch = 0; ch < p.nch; ++ch) { for (uint32_t i = tid; i < job.nsmpl; i += kT) { nz |= load_sample(st, ch, job.offset + i); } }
        return __syncthreads_or(nz) == 0;
    };

    if (p.serial_streams) {
        /* ODD block size: every block's window keeps a sample of the previous call (lpc.c:260-264), so the calls of a
         * stream form one chain from its first block to its last.  A CTA walks a stream's blocks in order with `pbuf`
         * never reset in between -- the reference's calculator, replayed call for call; streams run side by side. */
        uint32_t owner = 0xffffffffu, seen = 0;
        for (uint32_t j = 0; j < p.num_jobs; ++j) {
            const Job job = p.jobs[j];
            if (job.flags & kJobFirstOfStream) { owner = seen % gridDim.x; seen++; }
            if (owner != blockIdx.x) { continue; }
            const StreamDev st = p.streams[job.stream];
            const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;
            __syncthreads();
            if (job.flags & kJobFirstOfStream) { for (uint32_t i = tid; i < pbuf_len; i += kT) { pbuf[i] = 0.0; } }
            if (job.nsmpl <= P || block_is_silent(job, st)) { flags_only(job, st, lshift, j); continue; }
            for (uint32_t cand = 0; cand < p.ncand; ++cand) {
                const uint32_t idx = j * p.ncand + cand;
                double *g = p.lags + (size_t)(idx >> 5) * p.lag_stride * 32u + (idx & 31u);
                step(job, st, cand, lshift, p.cand + idx, g, 32u, true);
            }
        }
        return;
    }
    for (uint32_t j = blockIdx.x; j < p.num_jobs; j += gridDim.x) {
        const Job job = p.jobs[j];
        const StreamDev st = p.streams[job.stream];
        const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;
        const uint32_t n = job.nsmpl;
        __syncthreads();
        if (n <= P) { flags_only(job, st, lshift, j); continue; }
        for (uint32_t i = tid; i < pbuf_len; i += kT) { pbuf[i] = 0.0; }
        /* with fixed blocks only a stream's last block can be odd or that short (the block size itself is even and >= 263) */
        const bool chain = p.replay_tails && ((n & 1u) || (kLtp && ceil_pow2_u32(n) < 263u));
        if (chain) {
            uint32_t pred = p.group_first + j; bool found = false;
            while (!found && !(p.jobs_all[pred].flags & kJobFirstOfStream)) {
                pred--;
                const Job pj = p.jobs_all[pred];
                found = (pj.nsmpl > P) && !block_is_silent(pj, st);
            }
            if (found) { const Job pj = p.jobs_all[pred]; step(pj, st, p.ncand - 1u, lshift, &scratch_out, pbuf + p.fft_max, 1u, true); }     /* lags into pbuf's tail: a dump area */
        }
        for (uint32_t cand = 0; cand < p.ncand; ++cand) {
            const uint32_t idx = j * p.ncand + cand;
            double *g = p.lags + (size_t)(idx >> 5) * p.lag_stride * 32u + (idx & 31u);
            step(job, st, cand, lshift, p.cand + idx, g, 32u, chain);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * lpc_kernel: one THREAD per candidate (32 candidates per CTA).  Ridge, Levinson-Durbin for all
 * orders, order choice, coefficient quantisation (lpc.c:379-441, 483-493, 1341-1405;
 * srla_encoder.c:934-957, 1104-1108).
 * The recursion is sequential in the order and its reflection numerators are sequential sums in the
 * reference's order, so the work of one candidate is a dependent chain: it gets one lane, and the
 * lags / coefficient vectors of the 32 candidates of a warp are interleaved in shared memory
 * ([index][lane]) so every access is conflict free.  The coefficient vector is updated in place in
 * symmetric pairs (new[i], new[k+1-i] depend only on prev[i], prev[k+1-i]).
 * ---------------------------------------------------------------------------------------------- */
/* estimated size of a block coded with the error variance of one order (srla_encoder.c:938-950, with the
 * window-compensated variance of lpc.c:493) */
__device__ __forceinline__ double order_bits(double err_k, uint32_t k, double gain, uint32_t n, uint32_t bps)
{
    const double ev = err_k * gain;
    const double mean_abs = 2.0 * sqrt(ev / 2.0);
    double bits = geometric_entropy(mean_abs, bps) * (double)n;
    bits += (double)(8u * k);
    return bits;
}

#define LPC_R(i) R[(size_t)(i) * 32u + lane]
#define LPC_A(i) A[(size_t)(i) * 32u + lane]
/* step k of the recursion's coefficient update: new[i] = prev[i] + refl * prev[k+1-i] (lpc.c:432-435) in symmetric
 * pairs; four pairs are loaded before any is stored so the shared-memory latencies overlap */
__device__ __forceinline__ void levinson_update(double *A, const uint32_t lane, const uint32_t k, const double refl)
{
    LPC_A(k + 1u) = 0.0;
    const uint32_t npairs = (k + 2u) >> 1;
    uint32_t j = 0;
    for (; j + 4u <= npairs; j += 4u) {
        const double a0 = LPC_A(j), a1 = LPC_A(j + 1u), a2 = LPC_A(j + 2u), a3 = LPC_A(j + 3u);
        const double b0 = LPC_A(k + 1u - j), b1 = LPC_A(k - j), b2 = LPC_A(k - 1u - j), b3 = LPC_A(k - 2u - j);
        LPC_A(j) = a0 + refl * b0; LPC_A(j + 1u) = a1 + refl * b1; LPC_A(j + 2u) = a2 + refl * b2; LPC_A(j + 3u) = a3 + refl * b3;
        LPC_A(k + 1u - j) = b0 + refl * a0; LPC_A(k - j) = b1 + refl * a1; LPC_A(k - 1u - j) = b2 + refl * a2; LPC_A(k - 2u - j) = b3 + refl * a3;
    }
    for (; j < npairs; ++j) {
        const double t1 = LPC_A(j), t2 = LPC_A(k + 1u - j);
        LPC_A(j) = t1 + refl * t2;
        LPC_A(k + 1u - j) = t2 + refl * t1;
    }
    if (((k + 1u) & 1u) == 0u) { const uint32_t mid = (k + 1u) >> 1; const double t = LPC_A(mid); LPC_A(mid) = t + refl * t; }
}

/* ------------------------------------------------------------------------------------------------
 * lpc_levinson_kernel: one THREAD per candidate (32 candidates per CTA).  Ridge and the Levinson-Durbin
 * recursion for all orders (lpc.c:379-441, 483-493).  The recursion is sequential in the order and its
 * reflection numerators are sequential sums in the reference's order, so the work of one candidate is a
 * dependent chain: it gets one lane, and the lags / coefficient vectors of the 32 candidates of a warp are
 * interleaved in shared memory ([index][lane]) so every access is conflict free.  The kernel keeps ONLY what the
 * chain needs: it records every order's reflection coefficient and error variance ([group][2][P+2][32] doubles)
 * and leaves the order choice (64 independent entropy estimates per candidate -- `log`, `sqrt`, divisions) and
 * the rebuild of the chosen order's coefficients to lpc_select_kernel, which has the parallelism for them.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(32) lpc_levinson_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t P = p.max_order;
    const uint32_t lane = threadIdx.x;
    const uint32_t total = p.num_jobs * p.ncand;
    const uint32_t first = blockIdx.x * 32u;
    double *R = reinterpret_cast<double *>(smem);            /* [P + 2][32] */
    double *A = R + (size_t)(P + 2u) * 32u;                  /* [P + 3][32] */
    double *g_refl = p.lpc_state + (size_t)blockIdx.x * 2u * (P + 2u) * 32u;     /* [P + 2][32]: a1, then the reflection of every step */
    double *g_err = g_refl + (size_t)(P + 2u) * 32u;                            /* [P + 2][32]: error variance of every order */
    /* the front kernel stores the lags of 32 consecutive candidates interleaved ([lag][candidate % 32]):
     * coalesced loads, conflict-free stores */
    {
        const double *g = p.lags + (size_t)blockIdx.x * p.lag_stride * 32u;
        for (uint32_t i = 0; i <= P; ++i) { R[(size_t)i * 32u + lane] = g[(size_t)i * 32u + lane]; }
    }
    __syncwarp();
    const uint32_t idx = first + lane;
    if (idx >= total) { return; }
    const uint32_t job_id = idx / p.ncand;
    const uint32_t n = p.jobs[job_id].nsmpl;
    const CandOut *out = p.cand + idx;
    if (n <= P || out->status != 0u) { return; }
    const double gain = p.jobs[job_id].welch_gain;
    CandDiag *dg = p.diag ? p.diag + idx : nullptr;
    LPC_R(0) = LPC_R(0) * (1.0 + 1e-5);                      /* ridge, lpc.c:483 */
    if (dg) { for (uint32_t i = 0; i <= P; ++i) { dg->autocorr[i] = LPC_R(i); } }
    const double r0 = LPC_R(0), r1 = LPC_R(1);
    if (fabs(r0) < (double)FLT_EPSILON) {
        /* lpc.c:399-407: all coefficient vectors zero, every variance r[0] */
        for (uint32_t k = 0; k <= P; ++k) { g_refl[(size_t)k * 32u + lane] = 0.0; g_err[(size_t)k * 32u + lane] = r0; }
        if (dg) { for (uint32_t i = 0; i <= P; ++i) { dg->error_vars[i] = r0 * gain; } }
        return;
    }
    double e = r0;
    const double a1 = -r1 / r0;
    LPC_A(0) = 1.0; LPC_A(1) = a1;
    e = e + r1 * a1;
    g_refl[lane] = a1; g_err[lane] = r0; g_err[32u + lane] = e;
    if (dg) { dg->error_vars[0] = r0 * gain; dg->error_vars[1] = e * gain; }
    for (uint32_t k = 1; k < P; ++k) {
        double acc = 0.0;
        uint32_t i = 0;
        for (; i + 4u <= k + 1u; i += 4u) {
            const double m0 = LPC_A(i) * LPC_R(k + 1u - i), m1 = LPC_A(i + 1u) * LPC_R(k - i);
            const double m2 = LPC_A(i + 2u) * LPC_R(k - 1u - i), m3 = LPC_A(i + 3u) * LPC_R(k - 2u - i);
            acc += m0; acc += m1; acc += m2; acc += m3;
        }
        for (; i <= k; ++i) { acc += LPC_A(i) * LPC_R(k + 1u - i); }
        const double refl = acc / (-e);
        e = e * (1.0 - refl * refl);
        levinson_update(A, lane, k, refl);
        g_refl[(size_t)k * 32u + lane] = refl; g_err[(size_t)(k + 1u) * 32u + lane] = e;
        if (dg) { dg->error_vars[k + 1u] = e * gain; }
    }
}

/* 8-bit quantisation of the coefficient vector a(0) .. a(order-1) with error feedback from the tail
 * (lpc.c:1341-1405), stored reversed for the FIR (srla_encoder.c:1104-1108); returns the right shift */
template <typename Get>
__device__ __forceinline__ uint32_t quantise_coefficients(Get a, uint32_t order, int16_t *coef)
{
    double peak = 0.0;
    for (uint32_t i = 0; i < order; ++i) { const double v = fabs(a(i)); if (peak < v) { peak = v; } }
    if (peak <= 0.0078125) {
        for (uint32_t i = 0; i < order; ++i) { coef[i] = 0; }
        return 8u;
    }
    int exponent;
    (void)frexp(peak, &exponent);
    uint32_t rshift = (uint32_t)(7 - exponent);
    if (rshift >= 16u) { rshift = 15u; }
    const double scale = (double)(1u << rshift);
    double carry = 0.0;
    for (int i = (int)order - 1; i >= 0; --i) {
        carry += a((uint32_t)i) * scale;
        int32_t v = (int32_t)round_half_away(carry);
        if (v >= 128) { v = 127; } else if (v < -128) { v = -128; }
        carry -= (double)v;
        /* FIR order: coef[j] multiplies x[n - order + j]  =>  quantised a[order-1-j] */
        coef[order - 1u - (uint32_t)i] = (int16_t)v;
    }
    return rshift;
}

/* ------------------------------------------------------------------------------------------------
 * lpc_select_kernel: one CTA (128 threads) per 32 candidates.  Phase 1, all four warps: the estimated size of
 * every order (thread = candidate x order mod 4), first minimum (srla_encoder.c:934-957).  Phase 2, warp 0, one
 * lane per candidate: the chosen order's coefficient vector is rebuilt from the recorded reflection coefficients
 * (the recursion's update steps replayed: the same operations on the same values as the reference's second run),
 * then quantised with error feedback from the tail (lpc.c:1341-1405) and reversed for the FIR
 * (srla_encoder.c:1104-1108).
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) lpc_select_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t P = p.max_order, bps = p.bps;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, part = tid >> 5;
    const uint32_t total = p.num_jobs * p.ncand;
    double *A = reinterpret_cast<double *>(smem);                               /* [P + 3][32] */
    double *best_bits = A + (size_t)(P + 3u) * 32u;                             /* [4][32] */
    uint32_t *best_arg = reinterpret_cast<uint32_t *>(best_bits + 4u * 32u);    /* [4][32] */
    const double *g_refl = p.lpc_state + (size_t)blockIdx.x * 2u * (P + 2u) * 32u;
    const double *g_err = g_refl + (size_t)(P + 2u) * 32u;
    const uint32_t idx = blockIdx.x * 32u + lane;
    bool live = idx < total;
    uint32_t n = 0; double gain = 0.0;
    CandOut *out = p.cand + (live ? idx : 0u);
    if (live) {
        const uint32_t job_id = idx / p.ncand;
        n = p.jobs[job_id].nsmpl; gain = p.jobs[job_id].welch_gain;
        live = (n > P) && (out->status == 0u);
    }
    /* phase 1: orders part + 1, part + 5, ... ; strict comparison in ascending order = the reference's first minimum.
     * The variances of eight orders are fetched before any is evaluated: one memory latency per eight estimates. */
    {
        double best = (double)FLT_MAX; uint32_t arg = 0u;
        if (live) {
            for (uint32_t k0 = 1u + part; k0 <= P; k0 += 32u) {
                double ev[8];
                #pragma unroll
                for (int t = 0; t < 8; ++t) { const uint32_t k = k0 + 4u * (uint32_t)t; ev[t] = (k <= P) ? g_err[(size_t)k * 32u + lane] : 0.0; }
                #pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t k = k0 + 4u * (uint32_t)t;
                    if (k <= P) {
                        const double bits = order_bits(ev[t], k, gain, n, bps);
                        if (best > bits) { best = bits; arg = k; }
                    }
                }
            }
        }
        best_bits[part * 32u + lane] = best; best_arg[part * 32u + lane] = arg;
    }
    __syncthreads();
    if (part != 0u || !live) { return; }
    uint32_t order = 0u;
    {
        double best = (double)FLT_MAX;
        for (uint32_t q = 0; q < 4u; ++q) {
            const double b = best_bits[q * 32u + lane]; const uint32_t a = best_arg[q * 32u + lane];
            if (a != 0u && (best > b || (best == b && a < order))) { best = b; order = a; }
        }
    }
    CandDiag *dg = p.diag ? p.diag + idx : nullptr;
    uint32_t rshift = 0;
    if (order > 0u) {
        /* phase 2: replay the update steps 1 .. order-1 (reflection coefficients fetched one step ahead) */
        LPC_A(0) = 1.0; LPC_A(1) = g_refl[lane];
        double next = (order > 1u) ? g_refl[32u + lane] : 0.0;
        for (uint32_t k = 1; k < order; ++k) {
            const double refl = next;
            if (k + 1u < order) { next = g_refl[(size_t)(k + 1u) * 32u + lane]; }
            levinson_update(A, lane, k, refl);
        }
        rshift = quantise_coefficients([&](uint32_t i) { return LPC_A(1u + i); }, order, out->coef);
        if (p.svr_iterations) { double *keep = p.svr_coef + (size_t)idx * P; for (uint32_t i = 0; i < order; ++i) { keep[i] = LPC_A(1u + i); } }
        if (dg) { for (uint32_t i = 0; i < order; ++i) { dg->lpc_double[i] = LPC_A(1u + i); } }
    }
    out->order = order; out->rshift = rshift;
}
#undef LPC_R
#undef LPC_A

/* ------------------------------------------------------------------------------------------------
 * svr_kernel (only with num_svr_filter_learning_iteration > 0): the reference refines the chosen
 * order's coefficients by iteratively re-weighted least squares with a soft-thresholded ("support vector")
 * residual before quantising them (LPC_CalculateCoefSVR, lpc.c:1036-1136; call site srla_encoder.c:1087-1101).
 * Persistent CTAs, one candidate at a time per CTA.  Every floating-point sum is taken in the reference's order;
 * the parallelism is across sums:
 *   covariance   cov[i][j] = sum_s x[s+i] x[s+j]  (lpc.c:988-1017): a thread owns eight neighbouring j of one i
 *   Cholesky     (lpc.c:573-602): column by column, the entries of a column in parallel; L(j,i) overwrites cov[i][j]
 *   iteration    residual per sample in parallel (taps in order), then |residual| summed by ONE thread in sample
 *                order while other threads accumulate the right-hand side r[i] (one i each, samples in order),
 *                then the two triangular solves (lpc.c:605-631) and the bookkeeping by one thread.
 * The objective (lpc.c:1020-1030) uses log() and pow(); it only ever enters comparisons.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ double svr_objective(double mean_abs)
{
    const double intmean = mean_abs * 65536.0;                       /* BITS_PER_SAMPLE is fixed to 16 there */
    const double rho = 1.0 / (1.0 + intmean);
    const double l2 = log(log(0.5127629514) / log(1.0 - rho)) * 1.4426950408889634;
    const uint32_t k2 = (uint32_t)((0.0 > l2) ? 0.0 : l2);
    const uint32_t k1 = k2 + 1u;
    const double k1factor = pow(1.0 - rho, (double)(1u << k1));
    const double k2factor = pow(1.0 - rho, (double)(1u << k2));
    return (1.0 + k1) * (1.0 - k1factor) + (1.0 + k2 + (1.0 / (1.0 - k2factor))) * k1factor;
}

__global__ void __launch_bounds__(kThreads) svr_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const SvrLayout L = make_svr_layout(p.nmax, p.max_order);
    const uint32_t P = p.max_order, tid = threadIdx.x;
    double *data = reinterpret_cast<double *>(smem + L.data_off) + 8;          /* 8 zeros in front, 24 behind */
    double *resid = reinterpret_cast<double *>(smem + L.resid_off);
    double *vec = reinterpret_cast<double *>(smem + L.vec_off);
    double *coef = vec, *best = vec + (P + 1u), *init = vec + 2u * (P + 1u), *delta = vec + 3u * (P + 1u),
           *rvec = vec + 4u * (P + 1u), *inv_diag = vec + 5u * (P + 1u), *scal = vec + 6u * (P + 1u);
    /* scal[0] mean-abs sum, scal[1] stop flag, scal[2] singular flag */
    double *M = p.svr_matrix + (size_t)blockIdx.x * P * P;
    const uint32_t total = p.num_jobs * p.ncand;

    for (uint32_t idx = blockIdx.x; idx < total; idx += gridDim.x) {
        const uint32_t job_id = idx / p.ncand, cand = idx % p.ncand;
        const Job job = p.jobs[job_id];
        const StreamDev st = p.streams[job.stream];
        const uint32_t n = job.nsmpl;
        CandOut *out = p.cand + idx;
        __syncthreads();                                             /* the previous candidate is done with shared memory */
        if (n <= P || out->status != 0u || out->order == 0u) { continue; }
        const uint32_t dim = out->order, ltp_period = out->ltp_period;
        const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;

        /* ---- the signal the LPC stage saw, normalised to [-1, 1) (srla_encoder.c:1060-1064) ---- */
        {
            int32_t *raw = reinterpret_cast<int32_t *>(resid);
            int32_t *sig = raw + round_up_u32(n, 4) + 4;
            (void)load_candidate<2>(st, job, p, cand, lshift, raw);
            __syncthreads();
            apply_preemphasis(raw, sig, n, out->pre_coef);
            __syncthreads();
            if (ltp_period > 0u) { apply_ltp(sig, raw, n, p.ltp_order, ltp_period, out->ltp_coef[0], out->ltp_coef[1], out->ltp_coef[2]); }
            for (uint32_t i = tid; i < n; i += kThreads) { data[i] = (double)sig[i] * p.unit; }
            if (tid < 8u) { data[-1 - (int)tid] = 0.0; }
            if (tid < 24u) { data[n + tid] = 0.0; }
            const double *start = p.svr_coef + (size_t)idx * P;
            for (uint32_t i = tid; i < dim; i += kThreads) { const double c0 = start[i]; coef[i] = c0; init[i] = c0; best[i] = c0; }
        }
        __syncthreads();

        /* ---- covariance: task = (row i, eight columns j0 .. j0+7), samples in order ---- */
        {
            const uint32_t chunks = (dim + 7u) >> 3, terms = n - dim;
            for (uint32_t t = tid; t < dim * chunks; t += kThreads) {
                const uint32_t i = t / chunks, j0 = (t - i * chunks) << 3;
                if (j0 + 7u < i) { continue; }
                double acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
                double w[8];
                #pragma unroll
                for (int k = 0; k < 7; ++k) { w[k] = data[j0 + (uint32_t)k]; }
                w[7] = 0.0;
                for (uint32_t s0 = 0; s0 < terms; s0 += 8u) {
                    #pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t sm = s0 + (uint32_t)u;
                        /* window w[(u+k) & 7] = data[sm + j0 + k]: one new element per sample */
                        w[(u + 7) & 7] = data[sm + j0 + 7u];
                        if (sm < terms) {
                            const double sv = data[sm + i];
                            #pragma unroll
                            for (int k = 0; k < 8; ++k) { acc[k] += sv * w[(u + k) & 7]; }
                        }
                    }
                }
                #pragma unroll
                for (int k = 0; k < 8; ++k) { const uint32_t j = j0 + (uint32_t)k; if (j >= i && j < dim) { M[(size_t)i * P + j] = acc[k]; } }
            }
        }
        __syncthreads();
        /* ---- ridge (lpc.c:1066-1068) and Cholesky; L(j,i), j > i, takes the place of cov[i][j] ---- */
        if (tid == 0u) { scal[2] = 0.0; }
        for (uint32_t i = tid; i < dim; i += kThreads) { M[(size_t)i * P + i] *= (1.0 + 1e-5); }
        __syncthreads();
        for (uint32_t i = 0; i < dim; ++i) {
            if (tid == 0u) {
                double sum = M[(size_t)i * P + i];
                for (int k = (int)i - 1; k >= 0; --k) { const double l = M[(size_t)k * P + i]; sum -= l * l; }
                if (sum <= 0.0) { scal[2] = 1.0; inv_diag[i] = 0.0; } else { inv_diag[i] = inv_sqrt_cr(sum); }
            }
            __syncthreads();
            if (scal[2] != 0.0) { break; }
            const double inv = inv_diag[i];
            for (uint32_t j = i + 1u + tid; j < dim; j += kThreads) {
                double sum = M[(size_t)i * P + j];
                for (int k = (int)i - 1; k >= 0; --k) { sum -= M[(size_t)k * P + i] * M[(size_t)k * P + j]; }
                M[(size_t)i * P + j] = sum * inv;
            }
            __syncthreads();
        }
        if (scal[2] != 0.0) {
            /* singular: all-zero input in theory; the reference clears the coefficients (lpc.c:1071-1076) */
            if (tid == 0u) { out->rshift = quantise_coefficients([&](uint32_t) { return 0.0; }, dim, out->coef); }
            continue;
        }

        /* ---- margins x iterations ---- */
        double min_obj = (double)FLT_MAX;                            /* thread 0's copy is the one that counts */
        for (int mi = 0; mi < 6; ++mi) {
            const double margin = (mi == 0) ? 0.0 : 1.0 / (double)(4096u >> (2 * (mi - 1)));   /* 0, 1/4096, 1/1024, 1/256, 1/64, 1/16 (srla_internal.c:27) */
            double prev_obj = (double)FLT_MAX;
            __syncthreads();
            for (uint32_t i = tid; i < dim; i += kThreads) { coef[i] = init[i]; }
            __syncthreads();
            for (uint32_t itr = 0; itr < p.svr_iterations; ++itr) {
                /* residual of the current iterate, taps in order (lpc.c:1097-1100) */
                for (uint32_t sm = dim + tid; sm < n; sm += kThreads) {
                    double r = data[sm];
                    for (uint32_t i = 0; i < dim; ++i) { r += coef[i] * data[sm - i - 1u]; }
                    resid[sm] = r;
                }
                __syncthreads();
                if (tid == 0u) {
                    double mabse = 0.0;
                    #pragma unroll 8
                    for (uint32_t sm = dim; sm < n; ++sm) { const double r = resid[sm]; mabse += (r > 0) ? r : -r; }
                    scal[0] = mabse;
                } else if (tid >= 32u) {
                    for (uint32_t i = tid - 32u; i < dim; i += kThreads - 32u) {
                        double acc = 0.0;
                        #pragma unroll 4
                        for (uint32_t sm = dim; sm < n; ++sm) {
                            const double r = resid[sm];
                            const double mag = ((r > 0) ? r : -r) - margin;
                            const double soft = (double)((r > 0) - (r < 0)) * ((mag > 0.0) ? mag : 0.0);      /* LPC_SOFT_THRESHOLD */
                            acc += soft * data[sm - i - 1u];
                        }
                        rvec[i] = acc;
                    }
                }
                __syncthreads();
                if (tid == 0u) {
                    const double obj = svr_objective(scal[0] / n);
                    /* cov * delta = r by the Cholesky factor (lpc.c:605-631) */
                    for (uint32_t i = 0; i < dim; ++i) {
                        double sum = rvec[i];
                        for (int j = (int)i - 1; j >= 0; --j) { sum -= M[(size_t)j * P + i] * delta[j]; }
                        delta[i] = sum * inv_diag[i];
                    }
                    for (int i = (int)dim - 1; i >= 0; --i) {
                        double sum = delta[i];
                        for (uint32_t j = (uint32_t)i + 1u; j < dim; ++j) { sum -= M[(size_t)i * P + j] * delta[j]; }
                        delta[i] = sum * inv_diag[i];
                    }
                    if (obj < min_obj) { for (uint32_t i = 0; i < dim; ++i) { best[i] = coef[i]; } min_obj = obj; }
                    const bool stop = (prev_obj < obj) || (fabs(prev_obj - obj) < 1e-8);
                    if (!stop) { for (uint32_t i = 0; i < dim; ++i) { coef[i] += delta[i]; } prev_obj = obj; }
                    scal[1] = stop ? 1.0 : 0.0;
                }
                __syncthreads();
                if (scal[1] != 0.0) { break; }
            }
        }
        __syncthreads();
        if (tid == 0u) {
            out->rshift = quantise_coefficients([&](uint32_t i) { return best[i]; }, dim, out->coef);
            if (p.diag) { CandDiag *dg = p.diag + idx; for (uint32_t i = 0; i < dim; ++i) { dg->lpc_double[i] = best[i]; } }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * residual_kernel: one CTA per (job, candidate).  Rebuilds the candidate signal, runs the int32
 * FIR (srla_lpc_predict.c:236-264), the residual coder search (srla_coder.c:349-483) and the
 * side-information accounting (srla_encoder.c:1122-1187).
 *
 * FIR.  The coefficients are 8-bit and, for 16-bit sources, the pre-emphasised signal almost always
 * fits 16 bits, so the products are taken two taps at a time with IDP.2A (dp2a: s16 x s8 pairs
 * accumulated in a wrapping int32 -- the same value mod 2^32 as the reference's int32 multiply-adds).
 * The signal is repacked once into 16-byte entries Z[j] = { (x[4j],x[4j+1]), (x[4j+2],x[4j+3]),
 * (x[4j+1],x[4j+2]), (x[4j+3],x[4j+4]) }: even outputs read the first two pair words, odd outputs the
 * last two, so no output needs a funnel shift.  A thread produces 8 consecutive outputs from a
 * sliding window of three entries: one LDS.128 per 16 IDP.2A.  Signals that do not fit 16 bits
 * (24-bit sources, loud side channels) take the int32 IMAD path.
 *
 * Rice search.  For blocks of 1024 * {1,2,3,4,8} samples every thread owns four finest partitions;
 * the mean pyramid is built in registers (warp shuffles above the thread level), so each thread
 * knows the coding parameter of every enclosing partition.  The bits of the nine coarse partition
 * orders are accumulated by parameter VALUE: the cost of the thread's samples is evaluated once per
 * distinct parameter in the warp's range and credited to every level using it.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t zpair_slot(uint32_t j) { return j ^ ((j >> 3) & 1u); }

/* coding parameter of a partition with mean m (srla_coder.c:262-324) */
__device__ __forceinline__ uint32_t coding_parameter(double m, uint32_t code_type, const double *rice_threshold)
{
    if (code_type == kCodeRice) {
        uint32_t k = 0;
        #pragma unroll 1
        for (int j = 1; j < 32; ++j) { if (m >= __ldg(rice_threshold + j)) { k = (uint32_t)j; } }   /* host-libm thresholds */
        return k;
    }
    /* floor(log2((uint32_t)max(1.0, g))) (srla_coder.c:305-309) is the binary exponent of g when g >= 1 and 0 below:
     * read it from the exponent field instead of converting (F2I.F64 issues at a quarter of the rate) */
    const double g = 0.66794162356 * (1.0 + m);
    const int e = ((__double2hiint(g) >> 20) & 0x7ff) - 1023;
    return (uint32_t)max(e, 0);
}

/* variable part of the code length of u with parameter k: recursive Rice max((u >> k) - 2, 0), Rice u >> k
 * (srla_coder.c:165-190, 333-347); the constant part (k + 2, k + 1) is added per partition */
template <bool kRice>
__device__ __forceinline__ uint32_t var_len(uint32_t u, uint32_t k)
{
    const uint32_t t = u >> k;
    return kRice ? t : (uint32_t)__viaddmax_s32_relu((int)t, -2, 0);
}

struct RiceResult { uint32_t code_type, porder, bits; uint32_t thread_bits, tb_valid; };

/* fast path: n = 1024 * NQ samples, 256 threads, thread t owns samples [4 NQ t, 4 NQ (t + 1)) */
template <int NQ>
__device__ __forceinline__ RiceResult rice_search_fast(const int32_t *res_s, unsigned char *scratch, uint32_t *red32,
                                                       CandOut *out, const double *rice_threshold)
{
    constexpr int S = 4 * NQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *warp_mean = reinterpret_cast<double *>(scratch);                              /* [8]   */
    unsigned long long *kshare = reinterpret_cast<unsigned long long *>(scratch + 64);    /* [256] */
    uint16_t *thread_bits_all = reinterpret_cast<uint16_t *>(scratch + 64 + 8 * kThreads);   /* [11][256], saturated */
    RiceResult rr; rr.porder = 0; rr.thread_bits = 0; rr.tb_valid = 0;

    uint32_t v[S];
    {
        const int4 *src = reinterpret_cast<const int4 *>(res_s + (size_t)tid * S);
        #pragma unroll
        for (int c = 0; c < NQ; ++c) {
            const int4 t = src[c];
            v[4 * c] = zigzag32(t.x); v[4 * c + 1] = zigzag32(t.y); v[4 * c + 2] = zigzag32(t.z); v[4 * c + 3] = zigzag32(t.w);
        }
    }
    /* finest partition means: exact integer sums / NQ (srla_coder.c:371-383) */
    double m10[4];
    uint32_t any = 0;
    #pragma unroll
    for (int f = 0; f < 4; ++f) {
        unsigned long long s = 0;
        #pragma unroll
        for (int i = 0; i < NQ; ++i) { s += v[f * NQ + i]; any |= v[f * NQ + i]; }
        m10[f] = (double)s / (double)NQ;
    }
    any = (uint32_t)__syncthreads_or((int)(any != 0u));
    if (!any) { rr.code_type = kCodeAllZero; rr.bits = 2u; return rr; }

    /* mean pyramid (srla_coder.c:385-389): levels 9, 8 in the thread, 7..3 across the warp, 2..0 across warps */
    double mean[11];
    const double m9a = (m10[0] + m10[1]) / 2.0, m9b = (m10[2] + m10[3]) / 2.0;
    mean[8] = (m9a + m9b) / 2.0;
    #pragma unroll
    for (int l = 7; l >= 3; --l) {
        const double mine = mean[l + 1], other = __shfl_xor_sync(0xffffffffu, mine, 1 << (7 - l));
        mean[l] = (mine + other) / 2.0;
    }
    if (lane == 0) { warp_mean[warp] = mean[3]; }
    __syncthreads();
    {
        const int w2 = warp & ~1, w4 = warp & ~3;
        mean[2] = (warp_mean[w2] + warp_mean[w2 + 1]) / 2.0;
        const double q0 = (warp_mean[w4] + warp_mean[w4 + 1]) / 2.0, q1 = (warp_mean[w4 + 2] + warp_mean[w4 + 3]) / 2.0;
        mean[1] = (q0 + q1) / 2.0;
        const double h0 = ((warp_mean[0] + warp_mean[1]) / 2.0 + (warp_mean[2] + warp_mean[3]) / 2.0) / 2.0;
        const double h1 = ((warp_mean[4] + warp_mean[5]) / 2.0 + (warp_mean[6] + warp_mean[7]) / 2.0) / 2.0;
        mean[0] = (h0 + h1) / 2.0;
    }
    const uint32_t code_type = (mean[0] < 2.0) ? kCodeRice : kCodeRecursiveRice;
    rr.code_type = code_type;

    /* coding parameters of every partition this thread lies in (levels 0..8) or owns (9, 10) */
    uint32_t k[9], k9[2], k10[4];
    #pragma unroll
    for (int l = 0; l <= 8; ++l) { k[l] = coding_parameter(mean[l], code_type, rice_threshold); }
    k9[0] = coding_parameter(m9a, code_type, rice_threshold); k9[1] = coding_parameter(m9b, code_type, rice_threshold);
    #pragma unroll
    for (int f = 0; f < 4; ++f) { k10[f] = coding_parameter(m10[f], code_type, rice_threshold); }
    {
        unsigned long long pk = 0;
        #pragma unroll
        for (int l = 0; l <= 8; ++l) { pk |= (unsigned long long)k[l] << (5 * l); }
        pk |= (unsigned long long)k9[1] << 45; pk |= (unsigned long long)k10[3] << 50;
        kshare[tid] = pk;
    }
    __syncthreads();
    const unsigned long long prev = kshare[(tid > 0) ? tid - 1 : 0];

    uint32_t acc[11];
    /* parameter fields (srla_coder.c:419-428, 445-456): 5 bits for the first partition, zig-zag delta + 1 otherwise;
     * a partition is accounted by its first thread */
    #pragma unroll
    for (int l = 0; l <= 8; ++l) {
        uint32_t bits = 0;
        if ((tid & ((1 << (8 - l)) - 1)) == 0) {
            const uint32_t kp = (uint32_t)(prev >> (5 * l)) & 31u;
            bits = (tid == 0) ? 5u : (zigzag32((int32_t)k[l] - (int32_t)kp) + 1u);
        }
        acc[l] = bits;
    }
    {
        const uint32_t kp9 = (uint32_t)(prev >> 45) & 31u, kp10 = (uint32_t)(prev >> 50) & 31u;
        acc[9] = ((tid == 0) ? 5u : (zigzag32((int32_t)k9[0] - (int32_t)kp9) + 1u)) + zigzag32((int32_t)k9[1] - (int32_t)k9[0]) + 1u;
        acc[10] = ((tid == 0) ? 5u : (zigzag32((int32_t)k10[0] - (int32_t)kp10) + 1u))
                + zigzag32((int32_t)k10[1] - (int32_t)k10[0]) + zigzag32((int32_t)k10[2] - (int32_t)k10[1])
                + zigzag32((int32_t)k10[3] - (int32_t)k10[2]) + 3u;
    }
    /* code bits */
    const uint32_t fixed = (code_type == kCodeRice) ? 1u : 2u;
    #pragma unroll
    for (int l = 0; l <= 8; ++l) { acc[l] += (uint32_t)S * (k[l] + fixed); }
    acc[9] += (uint32_t)(2 * NQ) * (k9[0] + k9[1] + 2u * fixed);
    acc[10] += (uint32_t)NQ * (k10[0] + k10[1] + k10[2] + k10[3] + 4u * fixed);
    uint32_t kmin = k[0], kmax = k[0];
    #pragma unroll
    for (int l = 1; l <= 8; ++l) { kmin = min(kmin, k[l]); kmax = max(kmax, k[l]); }
    const uint32_t wmin = __reduce_min_sync(0xffffffffu, kmin), wmax = __reduce_max_sync(0xffffffffu, kmax);
    if (code_type == kCodeRice) {
        for (uint32_t kk = wmin; kk <= wmax; ++kk) {
            uint32_t s = 0;
            #pragma unroll
            for (int i = 0; i < S; ++i) { s += var_len<true>(v[i], kk); }
            #pragma unroll
            for (int l = 0; l <= 8; ++l) { acc[l] += (k[l] == kk) ? s : 0u; }
        }
        #pragma unroll
        for (int i = 0; i < S; ++i) { acc[9] += var_len<true>(v[i], k9[i / (2 * NQ)]); acc[10] += var_len<true>(v[i], k10[i / NQ]); }
    } else {
        for (uint32_t kk = wmin; kk <= wmax; ++kk) {
            uint32_t s = 0;
            #pragma unroll
            for (int i = 0; i < S; ++i) { s += var_len<false>(v[i], kk); }
            #pragma unroll
            for (int l = 0; l <= 8; ++l) { acc[l] += (k[l] == kk) ? s : 0u; }
        }
        #pragma unroll
        for (int i = 0; i < S; ++i) { acc[9] += var_len<false>(v[i], k9[i / (2 * NQ)]); acc[10] += var_len<false>(v[i], k10[i / NQ]); }
    }
    #pragma unroll
    for (int l = 0; l <= 10; ++l) { thread_bits_all[l * kThreads + tid] = (uint16_t)min(acc[l], 0xffffu); }
    #pragma unroll
    for (int l = 0; l <= 10; ++l) { acc[l] = __reduce_add_sync(0xffffffffu, acc[l]); }
    if (lane == 0) {
        #pragma unroll
        for (int l = 0; l <= 10; ++l) { red32[warp * 12 + l] = acc[l]; }
    }
    __syncthreads();
    if (warp == 0) {
        /* lane l totals partition order l; first minimum (srla_coder.c:434, 463: strict <, lowest order wins ties) */
        uint32_t bits = 0xffffffffu;
        if (lane <= 10) {
            bits = (uint32_t)kLog2MaxParts;
            #pragma unroll
            for (int w = 0; w < kWarps; ++w) { bits += red32[w * 12 + lane]; }
        }
        const uint32_t lowest = __reduce_min_sync(0xffffffffu, bits);
        const uint32_t arg = (uint32_t)__ffs((int)__ballot_sync(0xffffffffu, bits == lowest)) - 1u;
        if (lane == 0) { red32[96] = lowest; red32[97] = arg; }
    }
    __syncthreads();
    const uint32_t best_bits = red32[96], best = red32[97];
    rr.porder = best; rr.bits = best_bits + 2u;
    /* ... and what THIS thread's samples cost at the chosen order (parameter fields of the partitions that start here
     * included): exactly what emit_kernel's sizing pass would compute for the same samples.  Every thread parked its
     * eleven per-order counts in shared memory before the reduction. */
    const uint32_t tb = thread_bits_all[best * kThreads + tid];
    if (best <= 8u) {
        if ((tid & ((1 << (8 - best)) - 1)) == 0) {
            uint32_t kb = k[0];
            #pragma unroll
            for (int l = 1; l <= 8; ++l) { if (best == (uint32_t)l) { kb = k[l]; } }
            out->kparam[tid >> (8 - best)] = (uint8_t)kb;
        }
    } else if (best == 9u) {
        out->kparam[2 * tid] = (uint8_t)k9[0]; out->kparam[2 * tid + 1] = (uint8_t)k9[1];
    } else {
        *reinterpret_cast<uchar4 *>(out->kparam + 4 * tid) = make_uchar4((unsigned char)k10[0], (unsigned char)k10[1], (unsigned char)k10[2], (unsigned char)k10[3]);
    }
    rr.thread_bits = tb; rr.tb_valid = 1u;
    return rr;
}

/* general path: any block length (srla_coder.c:349-483) */
__device__ RiceResult rice_search_general(const int32_t *res_s, const uint32_t n, unsigned char *scratch, uint32_t *red32,
                                          CandOut *out, const double *rice_threshold)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    RiceResult rr; rr.porder = 0; rr.thread_bits = 0; rr.tb_valid = 0;
    uint32_t max_porder = (uint32_t)(__ffs((int)n) - 1);
    if (max_porder > (uint32_t)kLog2MaxParts) { max_porder = (uint32_t)kLog2MaxParts; }
    const uint32_t nparts = 1u << max_porder, per = n >> max_porder;
    /* heap layout: level l occupies [(1<<l)-1, (2<<l)-1).  Each slot first holds the partition's mean
     * (double), then is overwritten by its packed parameters: bits [5j, 5j+5) = parameter of the
     * enclosing partition at level j, for every j <= l. */
    double *mean = reinterpret_cast<double *>(scratch);
    unsigned long long *pack = reinterpret_cast<unsigned long long *>(mean);
    /* finest partition means: exact integer sums / per (srla_coder.c:371-383) */
    uint32_t any = 0;
    if (per <= 32u) {
        for (uint32_t q = tid; q < nparts; q += kThreads) {
            unsigned long long s = 0;
            const int32_t *rp = res_s + q * per;
            for (uint32_t i = 0; i < per; ++i) { const uint32_t v = zigzag32(rp[i]); s += v; any |= v; }
            mean[(nparts - 1u) + q] = (double)s / (double)per;
        }
    } else {
        for (uint32_t q = warp; q < nparts; q += kWarps) {
            unsigned long long s = 0;
            const int32_t *rp = res_s + q * per;
            for (uint32_t i = lane; i < per; i += 32) { const uint32_t v = zigzag32(rp[i]); s += v; any |= v; }
            s = (unsigned long long)warp_sum_ll((long long)s);
            if (lane == 0) { mean[(nparts - 1u) + q] = (double)s / (double)per; }
        }
    }
    any = (uint32_t)__syncthreads_or((int)(any != 0u));
    if (!any) { rr.code_type = kCodeAllZero; rr.bits = 2u; return rr; }
    for (int lvl = (int)max_porder - 1; lvl >= 0; --lvl) {
        const uint32_t cnt = 1u << lvl, base = cnt - 1u, child = 2u * cnt - 1u;
        for (uint32_t q = tid; q < cnt; q += kThreads) { mean[base + q] = (mean[child + 2u * q] + mean[child + 2u * q + 1u]) / 2.0; }
        __syncthreads();
    }
    const uint32_t code_type = (mean[0] < 2.0) ? kCodeRice : kCodeRecursiveRice;
    rr.code_type = code_type;
    __syncthreads();
    /* coding parameter of every partition, top level first, packed with its ancestors' */
    for (uint32_t lvl = 0; lvl <= max_porder; ++lvl) {
        const uint32_t cnt = 1u << lvl, base = cnt - 1u;
        for (uint32_t q = tid; q < cnt; q += kThreads) {
            const uint32_t k = coding_parameter(mean[base + q], code_type, rice_threshold);
            const unsigned long long parent = lvl ? pack[(cnt >> 1) - 1u + (q >> 1)] : 0ull;
            pack[base + q] = parent | ((unsigned long long)(k & 31u) << (5u * lvl));
        }
        __syncthreads();
    }
    /* bits of every partition order */
    uint32_t acc[kLog2MaxParts + 1];
    #pragma unroll
    for (int l = 0; l <= kLog2MaxParts; ++l) { acc[l] = 0; }
    const unsigned long long *finest = pack + (nparts - 1u);
    if (per <= 32u) {
        for (uint32_t q = tid; q < nparts; q += kThreads) {
            const unsigned long long pk = finest[q];
            const int32_t *rp = res_s + q * per;
            for (uint32_t i = 0; i < per; ++i) {
                const uint32_t v = zigzag32(rp[i]);
                #pragma unroll
                for (int l = 0; l <= kLog2MaxParts; ++l) {
                    const uint32_t k = (uint32_t)(pk >> (5 * l)) & 31u;
                    acc[l] += (code_type == kCodeRice) ? (1u + k + var_len<true>(v, k)) : (2u + k + var_len<false>(v, k));
                }
            }
        }
    } else {
        for (uint32_t i = tid; i < n; i += kThreads) {
            const uint32_t v = zigzag32(res_s[i]);
            const unsigned long long pk = finest[i / per];
            #pragma unroll
            for (int l = 0; l <= kLog2MaxParts; ++l) {
                const uint32_t k = (uint32_t)(pk >> (5 * l)) & 31u;
                acc[l] += (code_type == kCodeRice) ? (1u + k + var_len<true>(v, k)) : (2u + k + var_len<false>(v, k));
            }
        }
    }
    /* levels above max_porder accumulated garbage (their packed fields are 0): ignored below.
     * parameter fields: 5 bits for the first partition, zig-zag delta + 1 for the others */
    #pragma unroll
    for (int l = 0; l <= kLog2MaxParts; ++l) {
        if ((uint32_t)l <= max_porder) {
            const uint32_t cnt = 1u << l, base = cnt - 1u;
            for (uint32_t q = tid; q < cnt; q += kThreads) {
                const uint32_t k = (uint32_t)(pack[base + q] >> (5 * l)) & 31u;
                const uint32_t kprev = q ? ((uint32_t)(pack[base + q - 1u] >> (5 * l)) & 31u) : 0u;
                acc[l] += (q == 0u) ? 5u : (zigzag32((int32_t)k - (int32_t)kprev) + 1u);
            }
        }
    }
    #pragma unroll
    for (int l = 0; l <= kLog2MaxParts; ++l) { acc[l] = warp_sum_u32(acc[l]); }
    if (lane == 0) {
        #pragma unroll
        for (int l = 0; l <= kLog2MaxParts; ++l) { red32[warp * 12 + l] = acc[l]; }
    }
    __syncthreads();
    uint32_t best_bits = 0xffffffffu, best = 0;
    for (uint32_t l = 0; l <= max_porder; ++l) {
        uint32_t bits = (uint32_t)kLog2MaxParts;
        #pragma unroll
        for (int w = 0; w < kWarps; ++w) { bits += red32[w * 12 + l]; }
        if (bits < best_bits) { best_bits = bits; best = l; }
    }
    rr.porder = best; rr.bits = best_bits + 2u;
    for (uint32_t q = tid; q < (1u << best); q += kThreads) {
        out->kparam[q] = (uint8_t)((pack[((1u << best) - 1u) + q] >> (5u * best)) & 31u);
    }
    return rr;
}

/* ---- FIR residual with IDP.2A (srla_lpc_predict.c:236-264), shared by residual_kernel and residual16_kernel ---- */
__device__ __forceinline__ void fir_residual_dp2a(const int4 *Z, const uint32_t zf, const int32_t *coef_b, const uint32_t n, const uint32_t order,
                                                  const uint32_t p4, const uint32_t rshift, const uint32_t half, int32_t *res_s, int32_t *res_g)
{
    const int tid = threadIdx.x;
    /* ---- FIR residual with IDP.2A, int32 wrapping.  Every group of 8 outputs that reaches past the warm-up (first
     * output >= order, or straddling it) runs the full padded filter: for an output i >= order the taps that fall in
     * front of the block carry zero coefficients (order is padded to p4 in FRONT), so they may read anything
     * addressable.  Warm-up outputs (i < order: first differences, srla_lpc_predict.c:251-254) are patched in
     * afterwards.  No output is ever computed by a serial loop: one slow thread would hold the whole CTA at the
     * next barrier. ---- */
    const uint32_t groups = (n + 7u) >> 3, nm = p4 >> 2;
    for (uint32_t g = tid; g < groups; g += kThreads) {
        const uint32_t n0 = g << 3;
        int32_t a0 = (int32_t)half, a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
        int4 za, zb;
        if (order > 0u && n0 + 8u > order) {
            const uint32_t j0 = zf + (n0 >> 2) - nm;                       /* >= 0: zf covers p4 / 4 entries */
            za = Z[zpair_slot(j0)]; zb = Z[zpair_slot(j0 + 1u)];
            for (uint32_t m = 0; m < nm; ++m) {
                const int4 zc = Z[zpair_slot(j0 + m + 2u)];
                const int32_t cf = coef_b[m];
                a0 = __dp2a_lo(za.x, cf, a0); a0 = __dp2a_hi(za.y, cf, a0);
                a1 = __dp2a_lo(za.z, cf, a1); a1 = __dp2a_hi(za.w, cf, a1);
                a2 = __dp2a_lo(za.y, cf, a2); a2 = __dp2a_hi(zb.x, cf, a2);
                a3 = __dp2a_lo(za.w, cf, a3); a3 = __dp2a_hi(zb.z, cf, a3);
                a4 = __dp2a_lo(zb.x, cf, a4); a4 = __dp2a_hi(zb.y, cf, a4);
                a5 = __dp2a_lo(zb.z, cf, a5); a5 = __dp2a_hi(zb.w, cf, a5);
                a6 = __dp2a_lo(zb.y, cf, a6); a6 = __dp2a_hi(zc.x, cf, a6);
                a7 = __dp2a_lo(zb.w, cf, a7); a7 = __dp2a_hi(zc.z, cf, a7);
                za = zb; zb = zc;
            }
        } else {
            za = Z[zpair_slot(zf + (n0 >> 2))]; zb = Z[zpair_slot(zf + (n0 >> 2) + 1u)];
        }
        /* za, zb now hold the entries of samples n0 .. n0+4 and n0+4 .. n0+8: the outputs' own samples */
        const int32_t x[8] = { (int32_t)(short)(za.x & 0xffff), za.x >> 16, (int32_t)(short)(za.y & 0xffff), za.y >> 16,
                               (int32_t)(short)(zb.x & 0xffff), zb.x >> 16, (int32_t)(short)(zb.y & 0xffff), zb.y >> 16 };
        int32_t r[8];
        if (order > 0u) {
            r[0] = (int32_t)((uint32_t)x[0] + (uint32_t)asr32(a0, rshift)); r[1] = (int32_t)((uint32_t)x[1] + (uint32_t)asr32(a1, rshift));
            r[2] = (int32_t)((uint32_t)x[2] + (uint32_t)asr32(a2, rshift)); r[3] = (int32_t)((uint32_t)x[3] + (uint32_t)asr32(a3, rshift));
            r[4] = (int32_t)((uint32_t)x[4] + (uint32_t)asr32(a4, rshift)); r[5] = (int32_t)((uint32_t)x[5] + (uint32_t)asr32(a5, rshift));
            r[6] = (int32_t)((uint32_t)x[6] + (uint32_t)asr32(a6, rshift)); r[7] = (int32_t)((uint32_t)x[7] + (uint32_t)asr32(a7, rshift));
            if (n0 < order) {
                /* warm-up outputs of this group; x[n0 - 1] is the first half of the previous entry's last pair */
                const int32_t before = (int32_t)(short)(Z[zpair_slot(zf + (n0 >> 2) - 1u)].w & 0xffff);
                #pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const uint32_t i = n0 + (uint32_t)t;
                    if (i < order) { r[t] = (i == 0u) ? x[0] : (int32_t)((uint32_t)x[t] - (uint32_t)(t ? x[t - 1] : before)); }
                }
            }
        } else {
            #pragma unroll
            for (int t = 0; t < 8; ++t) { r[t] = x[t]; }
        }
        *reinterpret_cast<int4 *>(res_s + n0) = make_int4(r[0], r[1], r[2], r[3]);
        if (res_g) { *reinterpret_cast<int4 *>(res_g + n0) = make_int4(r[0], r[1], r[2], r[3]); }
        if (n0 + 4u < n) {
            *reinterpret_cast<int4 *>(res_s + n0 + 4u) = make_int4(r[4], r[5], r[6], r[7]);
            if (res_g) { *reinterpret_cast<int4 *>(res_g + n0 + 4u) = make_int4(r[4], r[5], r[6], r[7]); }
        }
    }
}

/* ---- the same filter on an int32 signal (24-bit sources, signals that do not fit 16 bits) ---- */
__device__ __forceinline__ void fir_residual_imad(const int32_t *sig, const int32_t *coef_s, const uint32_t n, const uint32_t order,
                                                  const uint32_t p4, const uint32_t rshift, const uint32_t half, int32_t *res_s, int32_t *res_g)
{
    const int tid = threadIdx.x;
    if (order > 0u) {
        /* same scheme as above: padded filter for every group reaching past the warm-up, warm-up outputs patched in */
        const uint32_t groups = (n + 3u) >> 2;
        const int4 *coef4 = reinterpret_cast<const int4 *>(coef_s);
        for (uint32_t g = tid; g < groups; g += kThreads) {
            const uint32_t n0 = g << 2;
            uint32_t a0 = half, a1 = half, a2 = half, a3 = half;
            const int4 w_self = *reinterpret_cast<const int4 *>(sig + n0);
            if (n0 + 4u > order) {
                const int4 *xp = reinterpret_cast<const int4 *>(sig + n0) - (p4 >> 2);      /* may start in the front padding */
                int4 w0 = xp[0];
                for (uint32_t m = 0; m < (p4 >> 2); ++m) {
                    const int4 w1 = xp[m + 1u];
                    const int4 cf = coef4[m];
                    a0 += (uint32_t)cf.x * (uint32_t)w0.x + (uint32_t)cf.y * (uint32_t)w0.y + (uint32_t)cf.z * (uint32_t)w0.z + (uint32_t)cf.w * (uint32_t)w0.w;
                    a1 += (uint32_t)cf.x * (uint32_t)w0.y + (uint32_t)cf.y * (uint32_t)w0.z + (uint32_t)cf.z * (uint32_t)w0.w + (uint32_t)cf.w * (uint32_t)w1.x;
                    a2 += (uint32_t)cf.x * (uint32_t)w0.z + (uint32_t)cf.y * (uint32_t)w0.w + (uint32_t)cf.z * (uint32_t)w1.x + (uint32_t)cf.w * (uint32_t)w1.y;
                    a3 += (uint32_t)cf.x * (uint32_t)w0.w + (uint32_t)cf.y * (uint32_t)w1.x + (uint32_t)cf.z * (uint32_t)w1.y + (uint32_t)cf.w * (uint32_t)w1.z;
                    w0 = w1;
                }
            }
            int32_t r[4] = { (int32_t)((uint32_t)w_self.x + (uint32_t)asr32((int32_t)a0, rshift)), (int32_t)((uint32_t)w_self.y + (uint32_t)asr32((int32_t)a1, rshift)),
                             (int32_t)((uint32_t)w_self.z + (uint32_t)asr32((int32_t)a2, rshift)), (int32_t)((uint32_t)w_self.w + (uint32_t)asr32((int32_t)a3, rshift)) };
            if (n0 < order) {
                const int32_t x[4] = { w_self.x, w_self.y, w_self.z, w_self.w };
                const int32_t before = sig[n0 ? n0 - 1u : 0u];
                #pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t i = n0 + (uint32_t)t;
                    if (i < order) { r[t] = (i == 0u) ? x[0] : (int32_t)((uint32_t)x[t] - (uint32_t)(t ? x[t - 1] : before)); }
                }
            }
            *reinterpret_cast<int4 *>(res_s + n0) = make_int4(r[0], r[1], r[2], r[3]);
            if (res_g) { *reinterpret_cast<int4 *>(res_g + n0) = make_int4(r[0], r[1], r[2], r[3]); }
        }
    } else {
        for (uint32_t i = tid; i < n; i += kThreads) { const int32_t v = sig[i]; res_s[i] = v; if (res_g) { res_g[i] = v; } }
    }
}

/* ---- residual coder search, side-information bits, result record: the tail of both residual kernels ---- */
__device__ __forceinline__ void residual_finish(const LaunchParams &p, CandOut *out, const int32_t *res_s, const uint32_t n, unsigned char *scratch,
                                                uint32_t *red32, const int32_t *coef_s, const uint32_t p4, const uint32_t order, const uint32_t ltp_period)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t bps = p.bps;
    /* ---- residual coder search (srla_coder.c:349-483) ---- */
    RiceResult rr;
    switch (n) {
        case 1024u: rr = rice_search_fast<1>(res_s, scratch, red32, out, p.rice_threshold); break;
        case 2048u: rr = rice_search_fast<2>(res_s, scratch, red32, out, p.rice_threshold); break;
        case 3072u: rr = rice_search_fast<3>(res_s, scratch, red32, out, p.rice_threshold); break;
        case 4096u: rr = rice_search_fast<4>(res_s, scratch, red32, out, p.rice_threshold); break;
        case 8192u: rr = rice_search_fast<8>(res_s, scratch, red32, out, p.rice_threshold); break;
        default:    rr = rice_search_general(res_s, n, scratch, red32, out, p.rice_threshold); break;
    }
    const uint32_t code_type = rr.code_type, best_porder = rr.porder, residual_bits = rr.bits;
    if (rr.tb_valid) { out->thread_bits[tid] = (uint16_t)min(rr.thread_bits, 0xffffu); }

    /* ---- side-information bits (srla_encoder.c:1122-1187) and result ---- */
    if (tid < 32) {
        uint32_t plain_bits = 0, sum_bits = 0, bad = 0;
        const int32_t *cf = coef_s + (p4 - order);
        for (uint32_t i = lane; i < order; i += 32) {
            const int32_t c = cf[i];
            const uint32_t len = __ldg(p.huff_len + zigzag32(c));
            plain_bits += len;
            if (i == 0u) { sum_bits += len; }
            else {
                const uint32_t sym = zigzag32(c + cf[i - 1u]);
                if (sym >= 256u) { bad = 1; } else { sum_bits += __ldg(p.huff_len + 256 + sym); }
            }
        }
        plain_bits = warp_sum_u32(plain_bits); sum_bits = warp_sum_u32(sum_bits); bad = warp_sum_u32(bad);
        if (lane == 0) {
            uint32_t use_sum = 0, coef_bits = 0;
            if (order > 0u) {
                /* the reference's early exits are equivalent to: every symbol valid and the summed form strictly shorter */
                use_sum = (!bad && (order == 1u || sum_bits < plain_bits)) ? 1u : 0u;
                coef_bits = use_sum ? sum_bits : plain_bits;
            }
            uint32_t bits = residual_bits;
            bits += bps + 1u + 5u;
            bits += 8u + 4u + 1u;
            bits += coef_bits;
            bits += 1u;
            if (ltp_period > 0u) { bits += 1u + 8u + p.ltp_order * 6u; }
            out->use_sum = use_sum; out->coef_bits = coef_bits; out->tb_valid = rr.tb_valid;
            out->code_type = code_type; out->porder = best_porder;
            out->residual_bits = residual_bits; out->total_bits = bits;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * residual16_kernel: the residual stage for 16-bit PCM without LTP (configs 2, 4, 5) as PERSISTENT CTAs whose source
 * samples arrive by bulk asynchronous copies (cp.async.bulk, completion on an mbarrier) one work item ahead.
 *
 * While a CTA filters and searches candidate i, the int16 source rows of candidate i + gridDim.x are already on their way
 * into the other buffer, and the descriptors that copy needs (job, stream, the candidate's analysis record) were fetched by
 * cp.async during the phases before: nobody waits for a dependent chain of global loads at the start of an item, which cost
 * the one-CTA-per-candidate kernel a fifth of every CTA's life.  The rows are converted straight into the FIR's 16-bit pair
 * entries (offset shift, mid/side, pre-emphasis on the way), so the int32 candidate signal is never materialised; the int32
 * residual of the current item is written over its own source rows once they are packed.
 * Same arithmetic as residual_kernel (shared device functions); items a bulk copy cannot serve (rows that are not
 * 16-byte aligned or not a multiple of 8 samples, e.g. a stream's tail block) are loaded with plain loads.
 * ---------------------------------------------------------------------------------------------- */
#ifndef SRLA_R16_OCC
#define SRLA_R16_OCC 4
#endif
struct Resid16Desc {
    uint32_t n, skip, bulk, order, rshift, lshift, kind, two_rows;      /* kind 0: mid, 1: side, 2: a channel as it is */
    int32_t  pre_coef;
    uint32_t job_id, cand, stream, offset, pad[3];
};
static_assert(sizeof(Resid16Desc) == 64, "descriptor slot");

__device__ __forceinline__ uint32_t smem_addr(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
/* global -> shared bulk copy of `bytes` (a multiple of 16, both addresses 16-byte aligned); completes on `bar` */
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_16(void *dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_addr(dst)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_8(void *dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_addr(dst)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/* candidate sample from the (shifted) left / right samples (srla_utility.c:91-103) */
__device__ __forceinline__ int32_t cand16_value(int32_t l, int32_t r, uint32_t kind, uint32_t lshift)
{
    l = asr32(l, lshift);
    if (kind == 2u) { return l; }
    r = asr32(r, lshift);
    const int32_t side = (int32_t)((uint32_t)r - (uint32_t)l);
    return (kind == 1u) ? side : (int32_t)((uint32_t)l + (uint32_t)(side >> 1));
}

/* pre-emphasised candidate sample i of an item, 0 outside [0, n) (srla_utility.c:342-358, filter memory = first sample) */
__device__ __forceinline__ int32_t cand16_filtered(const short *row0, const short *row1, const Resid16Desc &d, int32_t i)
{
    if (i < 0 || (uint32_t)i >= d.n) { return 0; }
    const int32_t ip = i ? i - 1 : 0;
    const int32_t cur = cand16_value(row0[i], d.two_rows ? row1[i] : 0, d.kind, d.lshift);
    const int32_t prv = cand16_value(row0[ip], d.two_rows ? row1[ip] : 0, d.kind, d.lshift);
    return (int32_t)((uint32_t)cur - (uint32_t)((int32_t)((uint32_t)prv * (uint32_t)d.pre_coef) >> 4));
}

__global__ void __launch_bounds__(kThreads, SRLA_R16_OCC) residual16_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const Resid16Layout L = make_resid16_layout(p.nmax, p.max_order);
    unsigned char *scratch = smem + L.scratch_off;
    int32_t  *coef_s = reinterpret_cast<int32_t *>(smem + L.coef_off);
    int32_t  *coef_b = reinterpret_cast<int32_t *>(smem + L.coefb_off);
    uint32_t *red32  = reinterpret_cast<uint32_t *>(smem + L.red_off);
    unsigned char *stage_base = smem + L.stage_off;            /* two slots of 128 bytes: [0,16) job head, [16,80) candidate record head, [80,128) stream */
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + L.bar_off);
    const int tid = threadIdx.x;
    const uint32_t total = p.num_jobs * p.ncand;
    const uint32_t first_ch = (p.nch >= 2u) ? 2u : 0u;

    /* every thread: the item whose records are staged in slot b */
    auto describe = [&](uint32_t idx, uint32_t b) {
        const unsigned char *stage = stage_base + 128u * b;
        const uint32_t *jh = reinterpret_cast<const uint32_t *>(stage);                 /* stream, offset, nsmpl, flags */
        const int32_t *ch = reinterpret_cast<const int32_t *>(stage + 16);              /* CandOut head */
        const StreamDev *st = reinterpret_cast<const StreamDev *>(stage + 80);
        Resid16Desc d;
        d.job_id = idx / p.ncand; d.cand = idx - d.job_id * p.ncand;
        d.stream = jh[0]; d.offset = jh[1]; d.n = jh[2];
        d.pre_coef = ch[0]; d.order = (uint32_t)ch[2]; d.rshift = (uint32_t)ch[3];
        d.skip = (d.n <= p.max_order || ch[15] != 0) ? 1u : 0u;                        /* RAW block / failed analysis */
        d.lshift = p.use_fixed_lshift ? p.fixed_lshift : st->lshift;
        const bool ms = (p.nch >= 2u) && (d.cand < 2u);
        d.kind = ms ? d.cand : 2u; d.two_rows = ms ? 1u : 0u;
        const uint32_t chan = ms ? 0u : d.cand - first_ch;
        /* rows a bulk copy can serve: 16-byte aligned, a multiple of 8 samples */
        const unsigned long long a0 = reinterpret_cast<unsigned long long>(st->pcm) + 2ull * ((unsigned long long)chan * st->stride + d.offset);
        const unsigned long long a1 = reinterpret_cast<unsigned long long>(st->pcm) + 2ull * (st->stride + d.offset);       /* right channel (mid / side) */
        d.bulk = (!d.skip && (d.n & 7u) == 0u && (a0 & 15ull) == 0ull && (!ms || (a1 & 15ull) == 0ull)) ? 1u : 0u;
        d.pad[0] = (uint32_t)a0; d.pad[1] = (uint32_t)(a0 >> 32); d.pad[2] = (uint32_t)(a1 - a0);
        return d;
    };
    /* thread 0: the source rows of the item staged in slot b on their way into buffer b */
    auto fetch_rows = [&](uint32_t idx, uint32_t b) {
        const Resid16Desc d = describe(idx, b);
        if (d.bulk) {
            unsigned char *buf = smem + L.buf_off[b];
            const uint32_t bytes = 2u * d.n;
            const unsigned long long a0 = ((unsigned long long)d.pad[1] << 32) | d.pad[0];
            /* the buffer was last written through the generic proxy (the previous item's residual): order those accesses
             * before the asynchronous-proxy writes of the copy */
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar[b], d.two_rows ? 2u * bytes : bytes);
            bulk_copy_g2s(buf, reinterpret_cast<const void *>(a0), bytes, &bar[b]);
            if (d.two_rows) { bulk_copy_g2s(buf + L.row_bytes, reinterpret_cast<const void *>(a0 + d.pad[2]), bytes, &bar[b]); }
        }
    };
    /* thread 0: start fetching the records of item idx into slot b (job head and candidate record now; the stream once the
     * job is known) */
    auto stage_job_and_cand = [&](uint32_t idx, uint32_t b) {
        unsigned char *stage = stage_base + 128u * b;
        const uint32_t job_id = idx / p.ncand;
        const unsigned char *jsrc = reinterpret_cast<const unsigned char *>(p.jobs + job_id);
        cp_async_8(stage, jsrc); cp_async_8(stage + 8, jsrc + 8);
        const unsigned char *csrc = reinterpret_cast<const unsigned char *>(p.cand + idx);
        cp_async_16(stage + 16, csrc); cp_async_16(stage + 32, csrc + 16); cp_async_16(stage + 48, csrc + 32); cp_async_16(stage + 64, csrc + 48);
        cp_async_commit();
    };
    auto stage_stream = [&](uint32_t b) {
        unsigned char *stage = stage_base + 128u * b;
        cp_async_wait_all();
        const uint32_t stream = *reinterpret_cast<const volatile uint32_t *>(stage);
        const unsigned char *ssrc = reinterpret_cast<const unsigned char *>(p.streams + stream);
        cp_async_16(stage + 80, ssrc); cp_async_16(stage + 96, ssrc + 16); cp_async_16(stage + 112, ssrc + 32);
        cp_async_commit();
    };

    if (tid == 0) {
        mbar_init(&bar[0], 1u); mbar_init(&bar[1], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x < total) { stage_job_and_cand(blockIdx.x, 0u); stage_stream(0u); cp_async_wait_all(); fetch_rows(blockIdx.x, 0u); }
    }
    __syncthreads();

    uint32_t parity_bits = 0u;                                 /* bit b: phase parity of mbarrier b (no locally indexed array) */
    uint32_t it = 0;
    for (uint32_t idx = blockIdx.x; idx < total; idx += gridDim.x, ++it) {
        const uint32_t cur = it & 1u, nxt = cur ^ 1u;
        const uint32_t next_idx = idx + gridDim.x;
        const bool have_next = next_idx < total;
        const Resid16Desc d = describe(idx, cur);                                      /* every thread, from the staged records */
        if (tid == 0 && have_next) { stage_job_and_cand(next_idx, nxt); }            /* phase A of the next item's fetch */
        const uint32_t n = d.n, order = d.order, rshift = d.rshift;
        CandOut *out = p.cand + idx;
        unsigned char *buf = smem + L.buf_off[cur];
        const short *row0 = reinterpret_cast<const short *>(buf);
        const short *row1 = reinterpret_cast<const short *>(buf + L.row_bytes);
        int32_t *res_s = reinterpret_cast<int32_t *>(buf);
        const uint32_t p4 = round_up_u32(order, 4);
        int32_t my_coef = 0;                                                          /* loaded now, stored after the packing loop */
        if (!d.skip && (uint32_t)tid < p4 && (uint32_t)tid >= p4 - order) { my_coef = (int32_t)out->coef[(uint32_t)tid - (p4 - order)]; }

        if (!d.skip) {
            if (d.bulk) { mbar_wait(&bar[cur], (parity_bits >> cur) & 1u); }
            else {
                /* plain loads: a tail block, or rows the bulk copy cannot address */
                const StreamDev st = p.streams[d.stream];
                const uint32_t chan = d.two_rows ? 0u : d.cand - first_ch;
                short *w0 = reinterpret_cast<short *>(buf), *w1 = reinterpret_cast<short *>(buf + L.row_bytes);
                for (uint32_t i = tid; i < n; i += kThreads) {
                    w0[i] = (short)load_sample(st, chan, d.offset + i);
                    if (d.two_rows) { w1[i] = (short)load_sample(st, 1u, d.offset + i); }
                }
                __syncthreads();
            }
        }
        if (d.bulk) { parity_bits ^= 1u << cur; }

        int32_t *res_g = p.residual ? p.residual + (size_t)idx * p.res_stride : nullptr;
        const uint32_t half = (rshift > 0u) ? (1u << (rshift - 1u)) : 0x80000000u;
        int4 *Z = reinterpret_cast<int4 *>(scratch);
        const uint32_t zf = resid_pair_front(p.max_order);
        int fits = 1;
        if (!d.skip) {
            /* ---- rows -> 16-bit pair entries, two entries (8 samples) per work item.  With a block that is a multiple of 8
             * samples every item takes the same straight-line path: the sample in front of the block is the first sample
             * itself (filter memory, srla_utility.c:342-358) and the sample behind it is zero -- both are selects, so no
             * warp diverges and no thread runs a longer path than its neighbours (they would all wait for it at the
             * barrier below). ---- */
            const uint32_t items = round_up_u32(n, 8) >> 3;
            const uint32_t pc = (uint32_t)d.pre_coef, kind = d.kind, lshift = d.lshift;
            if ((n & 7u) == 0u) {
                for (uint32_t q = tid; q < items; q += kThreads) {
                    int32_t x[9], c[10];
                    const int4 a = *reinterpret_cast<const int4 *>(row0 + 8u * q);
                    int32_t l[10] = { 0, (int32_t)(short)(a.x & 0xffff), a.x >> 16, (int32_t)(short)(a.y & 0xffff), a.y >> 16,
                                      (int32_t)(short)(a.z & 0xffff), a.z >> 16, (int32_t)(short)(a.w & 0xffff), a.w >> 16, (int32_t)row0[8u * q + 8u] };
                    l[0] = q ? (int32_t)row0[8u * q - 1u] : l[1];
                    if (d.two_rows) {
                        const int4 b = *reinterpret_cast<const int4 *>(row1 + 8u * q);
                        int32_t r[10] = { 0, (int32_t)(short)(b.x & 0xffff), b.x >> 16, (int32_t)(short)(b.y & 0xffff), b.y >> 16,
                                          (int32_t)(short)(b.z & 0xffff), b.z >> 16, (int32_t)(short)(b.w & 0xffff), b.w >> 16, (int32_t)row1[8u * q + 8u] };
                        r[0] = q ? (int32_t)row1[8u * q - 1u] : r[1];
                        if (lshift != 0u) {
                            #pragma unroll
                            for (int t = 0; t < 10; ++t) { l[t] = asr32(l[t], lshift); r[t] = asr32(r[t], lshift); }
                        }
                        #pragma unroll
                        for (int t = 0; t < 10; ++t) {
                            const int32_t side = (int32_t)((uint32_t)r[t] - (uint32_t)l[t]);
                            c[t] = (kind == 1u) ? side : (int32_t)((uint32_t)l[t] + (uint32_t)(side >> 1));
                        }
                    } else {
                        #pragma unroll
                        for (int t = 0; t < 10; ++t) { c[t] = (lshift != 0u) ? asr32(l[t], lshift) : l[t]; }
                    }
                    #pragma unroll
                    for (int t = 0; t < 9; ++t) { x[t] = (int32_t)((uint32_t)c[t + 1] - (uint32_t)((int32_t)((uint32_t)c[t] * pc) >> 4)); }
                    if (8u * q + 8u >= n) { x[8] = 0; }
                    #pragma unroll
                    for (int t = 0; t < 8; ++t) { fits &= ((uint32_t)(x[t] + 32768) < 65536u) ? 1 : 0; }
                    Z[zpair_slot(zf + 2u * q)] = pack_pair_entry(x);
                    Z[zpair_slot(zf + 2u * q + 1u)] = pack_pair_entry(x + 4);
                    if (q == 0u) {
                        /* the entry in front of the block: zeros, then the first sample (the odd outputs' first pair) */
                        const int32_t f[5] = { 0, 0, 0, 0, x[0] };
                        Z[zpair_slot(zf - 1u)] = pack_pair_entry(f);
                    }
                }
            } else {
                for (uint32_t q = tid; q < items; q += kThreads) {
                    int32_t x[9];
                    #pragma unroll
                    for (int t = 0; t < 9; ++t) { x[t] = cand16_filtered(row0, row1, d, (int32_t)(8u * q) + t); }
                    #pragma unroll
                    for (int t = 0; t < 8; ++t) { fits &= ((uint32_t)(x[t] + 32768) < 65536u) ? 1 : 0; }
                    Z[zpair_slot(zf + 2u * q)] = pack_pair_entry(x);
                    Z[zpair_slot(zf + 2u * q + 1u)] = pack_pair_entry(x + 4);
                    if (q == 0u) { const int32_t f[5] = { 0, 0, 0, 0, x[0] }; Z[zpair_slot(zf - 1u)] = pack_pair_entry(f); }
                }
            }
            if ((uint32_t)tid < p4) {
                coef_s[tid] = my_coef;
                /* the same coefficients as bytes, four taps per word: p4 is a multiple of 4, so the four lanes of a word are
                 * active together */
                const uint32_t lanes = __activemask();
                uint32_t w = (uint32_t)my_coef & 0xffu;
                w |= (__shfl_down_sync(lanes, (uint32_t)my_coef & 0xffu, 1) << 8);
                w |= (__shfl_down_sync(lanes, (uint32_t)my_coef & 0xffu, 2) << 16);
                w |= (__shfl_down_sync(lanes, (uint32_t)my_coef & 0xffu, 3) << 24);
                if ((tid & 3) == 0) { coef_b[tid >> 2] = (int32_t)w; }
            }
        }
        if (tid == 0 && have_next) { stage_stream(nxt); }                             /* phase B */
        const bool zmode = __syncthreads_and(fits) != 0;

        if (!d.skip) {
            if (zmode) {
                fir_residual_dp2a(Z, zf, coef_b, n, order, p4, rshift, half, res_s, res_g);
            } else {
                /* a sample of the filtered signal does not fit 16 bits (a loud side channel): int32 signal in the scratch area.
                 * The entries are dead; every thread keeps its samples in registers across the barrier that retires them. */
                int32_t *sig = reinterpret_cast<int32_t *>(scratch) + resid_front_pad(p.max_order);
                const uint32_t lo = resid_front_pad(p.max_order), hi = round_up_u32(n, 4) + 12u;
                __syncthreads();
                for (int32_t i = -(int32_t)lo + tid; i < (int32_t)hi; i += kThreads) { sig[i] = cand16_filtered(row0, row1, d, i); }
                __syncthreads();
                fir_residual_imad(sig, coef_s, n, order, p4, rshift, half, res_s, res_g);
            }
        }
        if (tid == 0 && have_next) { cp_async_wait_all(); fetch_rows(next_idx, nxt); }   /* phase C: buffer nxt held the previous item's residual */
        __syncthreads();
        if (!d.skip) { residual_finish(p, out, res_s, n, scratch, red32, coef_s, p4, order, 0u); }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------------
 * front16_kernel: front_kernel for 16-bit PCM without LTP and transforms of at most 4096 points (configs 2, 4, 5), with
 * one CTA per JOB instead of one per candidate.  What front_kernel does per candidate with its own loads happens once per
 * job here:
 *   - the int16 rows of the job's channels arrive in shared memory by cp.async.bulk (completion on an mbarrier) instead of
 *     every candidate CTA loading and unpacking its channel(s) again: 2 rows per stereo block instead of 6, and no int32
 *     copy of the candidate in shared memory;
 *   - ONE pass over the rows gives the exact integer sums r0, r1 (srla_utility.c:214-257) of all candidates;
 *   - the mid/side samples are formed on the fly from the rows where the window pass reads them (srla_utility.c:91-103).
 * Same transform code and therefore the same lags, bit for bit, as front_kernel.  Measured on config 2: 1.25 -> 1.01 ms.
 * What did NOT pay here (all measured): persistent CTAs that prefetch the next job's rows (1.29 ms: the loop state costs
 * registers in the pass functions, whose spills then compete with the twiddle tables for L1 -- 12 % instead of 4 % of the
 * twiddle sectors missed); a table of the Welch weights instead of evaluating them (1.06 ms: 7 % fewer FP64 instructions,
 * but 16 KB more in L1).
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ int32_t lds_s16(uint32_t addr) { int32_t v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

struct RowSource {
    static constexpr uint32_t kFixedM = 2048u, kFixedThreads = 128u;      /* blocks of 4096 samples, front16_kernel's CTA */
    uint32_t row0, row1;          /* SHARED-space addresses of the staged rows (offset shift applied, row[-1] == row[0]); row1 only for mid / side */
    uint32_t kind;                /* 0: mid, 1: side, 2: the channel in row0 as it is */
    uint32_t n, half_n, pc;
    double div, dn1;
    bool full;                    /* n is the transform size and at least 32 */
    /* candidate samples i - 1, i, i + 1 (i even) */
    __device__ __forceinline__ void cand3(const uint32_t i, int32_t &cm, int32_t &c0, int32_t &c1) const
    {
        const uint32_t a = lds_u32(row0 + 2u * i);
        cm = lds_s16(row0 + 2u * i - 2u); c0 = (int32_t)(short)(a & 0xffffu); c1 = (int32_t)a >> 16;
        if (kind != 2u) {
            const uint32_t b = lds_u32(row1 + 2u * i);
            const int32_t sm = lds_s16(row1 + 2u * i - 2u) - cm, s0 = (int32_t)(short)(b & 0xffffu) - c0, s1 = ((int32_t)b >> 16) - c1;
            if (kind == 1u) { cm = sm; c0 = s0; c1 = s1; } else { cm += sm >> 1; c0 += s0 >> 1; c1 += s1 >> 1; }
        }
    }
    __device__ __forceinline__ int32_t emphasised(const int32_t cur, const int32_t prv) const
    {
        return (int32_t)((uint32_t)cur - (uint32_t)((int32_t)((uint32_t)prv * pc) >> 4));        /* srla_utility.c:342-358 */
    }
    __device__ __forceinline__ double one(const uint32_t i, const int32_t cur, const int32_t prv) const
    {
        if (i >= n) { return 0.0; }
        const uint32_t s = (i < half_n) ? i : (n - 1u - i);
        const double ds = int_to_double((int32_t)s);
        const double w = div * ds * (dn1 - ds);
        return int_to_double(emphasised(cur, prv)) * w;
    }
    __device__ __forceinline__ double2 element(const uint32_t e) const
    {
        const uint32_t i = 2u * e;
        if (i >= n) { return make_double2(0.0, 0.0); }
        int32_t cm, c0, c1;
        cand3(i, cm, c0, c1);
        return make_double2(one(i, c0, cm), one(i + 1u, c1, c0));
    }
    __device__ __forceinline__ void first_pass(double2 *x, uint32_t M, uint32_t nn, uint32_t lgs, const double2 *tw_a, const double2 *tw_b, uint32_t need) const;
    __device__ __forceinline__ void first_pass_fixed(double2 *x, const double2 *tw_a, const double2 *tw_b) const;
    /* the weights evaluated per sample from exact double arguments, as WindowSource::load_full does */
    __device__ __forceinline__ void load_full(double2 (&v)[4][4], const uint32_t tid, const uint32_t M) const
    {
        const double lo0 = int_to_double((int32_t)(2u * tid)), hi0 = int_to_double((int32_t)(n - 1u - 2u * tid));
        const double dq = int_to_double((int32_t)(M >> 3));
        #pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t e = tid + (uint32_t)jp * (M >> 4) + (uint32_t)j * (M >> 2);
                int32_t cm, c0, c1;
                cand3(2u * e, cm, c0, c1);
                const double off = dq * (double)(jp + 4 * j);
                const double ds0 = (j < 2) ? lo0 + off : hi0 - off, ds1 = (j < 2) ? ds0 + 1.0 : ds0 + -1.0;
                const double w0 = div * ds0 * (dn1 - ds0), w1 = div * ds1 * (dn1 - ds1);
                v[jp][j] = make_double2(int_to_double(emphasised(c0, cm)) * w0, int_to_double(emphasised(c1, c0)) * w1);
            }
        }
    }
};

/* RowSource's first pass as a real call whose arguments travel in registers (a struct by value would go through the stack) */
__device__ __noinline__ void fft_first_pass_rows(double2 *x, const uint32_t M, const uint32_t nn, const uint32_t lgs, const double2 *tw_a, const double2 *tw_b,
                                                 const uint32_t need, const uint32_t row0, const uint32_t row1, const uint32_t flags, const uint32_t n,
                                                 const uint32_t pc, const double div)
{
    RowSource ws;
    ws.row0 = row0; ws.row1 = row1; ws.kind = flags & 3u; ws.full = (flags >> 4) & 1u;
    ws.n = n; ws.half_n = n >> 1; ws.pc = pc; ws.div = div; ws.dn1 = (double)(int32_t)(n - 1u);
    fft_pair_pass_impl<RowSource, true>(x, M, nn, lgs, tw_a, tw_b, need, ws);
}
/* ... and with the transform size a literal (a full block: n == 2 kM) */
template <uint32_t kM>
__device__ __noinline__ void fft_first_pass_rows_fixed(double2 *x, const double2 *tw_a, const double2 *tw_b, const uint32_t row0, const uint32_t row1,
                                                       const uint32_t kind, const uint32_t pc, const double div)
{
    RowSource ws;
    ws.row0 = row0; ws.row1 = row1; ws.kind = kind; ws.full = true;
    ws.n = 2u * kM; ws.half_n = kM; ws.pc = pc; ws.div = div; ws.dn1 = (double)(int32_t)(2u * kM - 1u);
    fft_pair_pass_impl<RowSource, true, true>(x, kM, kM, 0u, tw_a, tw_b, kM, ws);
}
__device__ __forceinline__ void RowSource::first_pass_fixed(double2 *x, const double2 *tw_a, const double2 *tw_b) const
{
    fft_first_pass_rows_fixed<kFixedM>(x, tw_a, tw_b, row0, row1, kind, pc, div);
}
__device__ __forceinline__ void RowSource::first_pass(double2 *x, uint32_t M, uint32_t nn, uint32_t lgs, const double2 *tw_a, const double2 *tw_b, uint32_t need) const
{
    fft_first_pass_rows(x, M, nn, lgs, tw_a, tw_b, need, row0, row1, kind | (full ? 16u : 0u), n, pc, div);
}

/* exact warp sum of one int64 per lane whose magnitude stays below 2^50: low 24 bits and the (signed) rest separately */
__device__ __forceinline__ long long warp_sum_ll_redux(long long v)
{
    const uint32_t lo = __reduce_add_sync(0xffffffffu, (uint32_t)v & 0xffffffu);
    const int32_t hi = __reduce_add_sync(0xffffffffu, (int32_t)(v >> 24));
    return (long long)hi * (1ll << 24) + (long long)lo;
}

template <int kT>
__global__ void __launch_bounds__(kT, 3) front16_kernel(const __grid_constant__ LaunchParams p)
{
    static_assert(kT == 128, "four warps: the reduction scratch and the work split assume it");
    extern __shared__ __align__(16) unsigned char smem[];
    const Front16Layout &L = p.f16;                                                    /* computed by the host: the offsets cost no registers */
    double *region_d = reinterpret_cast<double *>(smem + L.region_off);
    long long *red = reinterpret_cast<long long *>(smem + L.red_off);                 /* [warp][10] */
    int32_t *sh_coef = reinterpret_cast<int32_t *>(smem + L.coef_off);                /* [ncand] */
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + L.bar_off);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t P = p.max_order, ncand = p.ncand, nch = p.nch;
    const uint32_t first_ch = (nch >= 2u) ? 2u : 0u;
    auto row_ptr = [&](uint32_t c) { return reinterpret_cast<short *>(smem + L.rows_off + c * L.row_bytes + 16u); };

    const uint32_t job_id = blockIdx.x;
    const Job job = p.jobs[job_id];
    const StreamDev &st = p.streams[job.stream];
    const uint32_t n = job.nsmpl;
    const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;
    const unsigned long long a0 = reinterpret_cast<unsigned long long>(st.pcm) + 2ull * job.offset;     /* channel 0's first sample */
    const unsigned long long row_step = 2ull * st.stride;
    CandOut *out0 = p.cand + (size_t)job_id * ncand;

    if (n <= P) {
        /* RAW block (srla_encoder.c:777-779): nothing to analyse */
        if ((uint32_t)tid < ncand) {
            CandOut *out = out0 + tid;
            int nz = 0;
            if ((uint32_t)tid >= first_ch) { const short *g = reinterpret_cast<const short *>(a0 + ((uint32_t)tid - first_ch) * row_step); for (uint32_t i = 0; i < n; ++i) { nz |= (int)__ldg(g + i); } }
            out->nonzero = (nz != 0); out->status = 0; out->order = 0; out->rshift = 0;
            out->ltp_period = 0; out->ltp_coef[0] = 0; out->ltp_coef[1] = 0; out->ltp_coef[2] = 0;
            out->total_bits = 0; out->residual_bits = 0; out->pre_coef = 0; out->pre_prev = 0;
        }
        return;
    }

    /* ---- the rows: bulk copies when they are 16-byte aligned and a multiple of 8 samples, else plain loads ---- */
    const bool bulk = (n & 7u) == 0u && (a0 & 15ull) == 0ull && (nch < 2u || (row_step & 15ull) == 0ull);
    if (bulk) {
        if (tid == 0) {
            mbar_init(bar, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar, nch * 2u * n);
            for (uint32_t c = 0; c < nch; ++c) { bulk_copy_g2s(row_ptr(c), reinterpret_cast<const void *>(a0 + c * row_step), 2u * n, bar); }
        }
        __syncthreads();                                   /* the barrier is initialised before anybody polls it */
        mbar_wait(bar, 0u);
    } else {
        for (uint32_t c = 0; c < nch; ++c) {
            const short *g = reinterpret_cast<const short *>(a0 + c * row_step);
            short *r = row_ptr(c);
            for (uint32_t i = tid; i < n; i += kT) { r[i] = __ldg(g + i); }
        }
        __syncthreads();
    }
    if (lshift != 0u) {
        /* the stream's common trailing zeros (srla_utility.c:177-203) come off once, in place */
        for (uint32_t c = 0; c < nch; ++c) { short *r = row_ptr(c); for (uint32_t i = tid; i < n; i += kT) { r[i] = (short)asr32((int32_t)r[i], lshift); } }
        __syncthreads();
    }
    if ((uint32_t)tid < nch) { short *r = row_ptr(tid); r[-1] = r[0]; }       /* filter memory of the pre-emphasis = the first sample */

    /* ---- r0 = sum x^2, r1 = sum x[i] x[i+1] of every candidate as exact integers; OR of the plain channels ---- */
    {
        const short *L0 = row_ptr(0), *R0 = row_ptr((nch >= 2u) ? 1u : 0u);
        long long s[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };          /* M r0 r1, S r0 r1, left r0 r1, right r0 r1 */
        uint32_t nzl = 0u, nzr = 0u;
        const uint32_t nq = n >> 3;
        for (uint32_t q = tid; q < nq; q += kT) {
            const int4 a = *reinterpret_cast<const int4 *>(L0 + 8u * q), b = *reinterpret_cast<const int4 *>(R0 + 8u * q);
            int32_t l[9] = { (int32_t)(short)(a.x & 0xffff), a.x >> 16, (int32_t)(short)(a.y & 0xffff), a.y >> 16,
                             (int32_t)(short)(a.z & 0xffff), a.z >> 16, (int32_t)(short)(a.w & 0xffff), a.w >> 16, (int32_t)L0[8u * q + 8u] };
            int32_t r[9] = { (int32_t)(short)(b.x & 0xffff), b.x >> 16, (int32_t)(short)(b.y & 0xffff), b.y >> 16,
                             (int32_t)(short)(b.z & 0xffff), b.z >> 16, (int32_t)(short)(b.w & 0xffff), b.w >> 16, (int32_t)R0[8u * q + 8u] };
            nzl |= (uint32_t)(a.x | a.y | a.z | a.w); nzr |= (uint32_t)(b.x | b.y | b.z | b.w);
            if (8u * q + 8u >= n) { l[8] = 0; r[8] = 0; }                     /* no successor: the last product is dropped */
            int32_t m[9], sd[9];
            #pragma unroll
            for (int t = 0; t < 9; ++t) { sd[t] = r[t] - l[t]; m[t] = l[t] + (sd[t] >> 1); }
            #pragma unroll
            for (int t = 0; t < 8; ++t) {
                s[0] += (long long)m[t] * m[t];   s[1] += (long long)m[t] * m[t + 1];
                s[2] += (long long)sd[t] * sd[t]; s[3] += (long long)sd[t] * sd[t + 1];
                s[4] += (long long)l[t] * l[t];   s[5] += (long long)l[t] * l[t + 1];
                s[6] += (long long)r[t] * r[t];   s[7] += (long long)r[t] * r[t + 1];
            }
        }
        for (uint32_t i = 8u * nq + tid; i < n; i += kT) {
            const int32_t l0 = L0[i], r0 = R0[i];
            const int32_t l1 = (i + 1u < n) ? (int32_t)L0[i + 1u] : 0, r1 = (i + 1u < n) ? (int32_t)R0[i + 1u] : 0;
            const int32_t sd0 = r0 - l0, sd1 = r1 - l1, m0 = l0 + (sd0 >> 1), m1 = (i + 1u < n) ? l1 + (sd1 >> 1) : 0;
            nzl |= (uint32_t)l0; nzr |= (uint32_t)r0;
            s[0] += (long long)m0 * m0;   s[1] += (long long)m0 * m1;
            s[2] += (long long)sd0 * sd0; s[3] += (long long)sd0 * sd1;
            s[4] += (long long)l0 * l0;   s[5] += (long long)l0 * l1;
            s[6] += (long long)r0 * r0;   s[7] += (long long)r0 * r1;
        }
        #pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] = warp_sum_ll_redux(s[k]); }
        nzl = __reduce_or_sync(0xffffffffu, nzl); nzr = __reduce_or_sync(0xffffffffu, nzr);
        if (lane == 0) {
            #pragma unroll
            for (int k = 0; k < 8; ++k) { red[warp * 10 + k] = s[k]; }
            red[warp * 10 + 8] = (long long)nzl; red[warp * 10 + 9] = (long long)nzr;
        }
    }
    __syncthreads();
    if ((uint32_t)tid < ncand) {
        /* pre-emphasis coefficient of candidate tid (srla_utility.c:214-257): the sums are below 2^53, so the reference's
         * double sums are exact as well and the quotient is the same double */
        const uint32_t slot = (nch >= 2u) ? (uint32_t)tid : 2u;                 /* mono: the channel's sums sit in the `left` slot */
        long long s0 = 0, s1 = 0, nz = 0;
        #pragma unroll
        for (int w = 0; w < kT / 32; ++w) { s0 += red[w * 10 + 2 * slot]; s1 += red[w * 10 + 2 * slot + 1]; nz |= red[w * 10 + 8 + (slot & 1u)]; }
        int32_t c = 0;
        if (s0 != 0) {
            const double v = ((double)s1 / (double)s0) * 16.0;
            c = (int32_t)round_half_away(v);
            if (c < -16) { c = -16; }
            if (c > 15) { c = 15; }
        }
        sh_coef[tid] = c;
        const bool plain = (uint32_t)tid >= first_ch;
        const int32_t l0 = row_ptr(0)[0], r0 = row_ptr((nch >= 2u) ? 1u : 0u)[0];
        const int32_t first = plain ? (int32_t)row_ptr((uint32_t)tid - first_ch)[0] : ((tid == 1) ? r0 - l0 : l0 + ((r0 - l0) >> 1));
        CandOut *out = out0 + tid;
        out->nonzero = (plain && nz != 0) ? 1u : 0u; out->status = 0; out->order = 0; out->rshift = 0;
        out->ltp_period = 0; out->ltp_coef[0] = 0; out->ltp_coef[1] = 0; out->ltp_coef[2] = 0;
        out->total_bits = 0; out->residual_bits = 0; out->pre_coef = c; out->pre_prev = first;
    }
    __syncthreads();

    /* ---- autocorrelation of every candidate (lpc.c:444-483) ---- */
    if (P == 0u) { return; }
    const uint32_t N = ceil_pow2_u32(n);
    for (uint32_t c = 0; c < ncand; ++c) {
        const bool ms = (nch >= 2u) && (c < 2u);
        RowSource ws;
        ws.kind = ms ? c : 2u;
        ws.row0 = smem_addr(ms ? row_ptr(0) : row_ptr(c - first_ch));
        ws.row1 = smem_addr(row_ptr((nch >= 2u) ? 1u : 0u));
        ws.n = n; ws.half_n = n >> 1; ws.pc = (uint32_t)sh_coef[c];
        ws.div = job.welch_div * p.unit; ws.dn1 = (double)(int32_t)(n - 1u);
        ws.full = (n == N) && (N >= 32u);
        const uint32_t idx = job_id * ncand + c;
        /* lags of 32 consecutive candidates are interleaved for the lpc kernel: [lag][candidate % 32] */
        double *g = p.lags + (size_t)(idx >> 5) * p.lag_stride * 32u + (idx & 31u);
        welch_autocorr_core<RowSource, false>(ws, n, region_d, g, 32u, P + 1u, job.ac_scale, p, nullptr);
    }
}

/* kBig: blocks beyond the shared-memory capacity -- the same code on a per-CTA scratch area in global memory, persistent CTAs */
template <bool kBig>
__device__ __forceinline__ void residual_item(const LaunchParams &p, unsigned char *smem, const uint32_t item)
{
    const ResidLayout L = make_resid_layout(p.nmax, p.max_order);
    int32_t  *region_i = reinterpret_cast<int32_t *>(smem + L.region_off);
    int32_t  *sig      = reinterpret_cast<int32_t *>(smem + L.sig_off) + resid_front_pad(p.max_order);
    int32_t  *coef_s   = reinterpret_cast<int32_t *>(smem + L.coef_off);
    int32_t  *coef_b   = reinterpret_cast<int32_t *>(smem + L.coefb_off);
    uint32_t *red32    = reinterpret_cast<uint32_t *>(smem + L.red_off);

    const int tid = threadIdx.x;
    const uint32_t job_id = item / p.ncand, cand = item % p.ncand;
    const Job job = p.jobs[job_id];
    const StreamDev st = p.streams[job.stream];
    const uint32_t n = job.nsmpl;
    const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;
    CandOut *out = p.cand + (size_t)job_id * p.ncand + cand;
    if (n <= p.max_order || out->status != 0u) { return; }
    const uint32_t order = out->order, rshift = out->rshift, ltp_period = out->ltp_period;
    const int32_t pre_coef = out->pre_coef;

    const uint32_t p4 = round_up_u32(order, 4);
    for (uint32_t i = tid; i < p4; i += kThreads) { coef_s[i] = (i < p4 - order) ? 0 : (int32_t)out->coef[i - (p4 - order)]; }
    int32_t *res_s = region_i;
    unsigned char *scratch = smem + L.region_off + round_up_u32(4u * round_up_u32(n, 4), 16);
    int32_t *res_g = p.residual ? p.residual + ((size_t)job_id * p.ncand + cand) * p.res_stride : nullptr;
    const uint32_t half = (rshift > 0u) ? (1u << (rshift - 1u)) : 0x80000000u;
    /* ---- the signal the FIR runs on, as 16-bit pair entries (entry j = samples 4j .. 4j+4 at Z[zpair_slot(j + zf)]) ---- */
    int4 *Z = reinterpret_cast<int4 *>(scratch);
    const uint32_t zf = resid_pair_front(p.max_order);
    const uint32_t zcount = round_up_u32(n, 8) >> 2;
    (void)load_candidate<2>(st, job, p, cand, lshift, region_i);        /* two quads (x 2 channels) in flight; four measured slower */
    __syncthreads();
    int fits = 1;
    if (ltp_period == 0u) {
        /* pre-emphasis (srla_utility.c:342-358, filter memory = first sample) folded into the packing: raw -> entries */
        const int32_t *raw = region_i;
        const uint32_t pc = (uint32_t)pre_coef;
        for (uint32_t jj = tid; jj < zcount + 1u; jj += kThreads) {
            const int32_t j = (int32_t)jj - 1;
            int32_t x[5];
            if (j >= 0 && 4u * (uint32_t)j + 4u < n) {
                const int4 q = *reinterpret_cast<const int4 *>(raw + 4 * j);
                const int32_t nxt = raw[4 * j + 4], prv = raw[j ? 4 * j - 1 : 0];
                x[0] = (int32_t)((uint32_t)q.x - (uint32_t)((int32_t)((uint32_t)prv * pc) >> 4));
                x[1] = (int32_t)((uint32_t)q.y - (uint32_t)((int32_t)((uint32_t)q.x * pc) >> 4));
                x[2] = (int32_t)((uint32_t)q.z - (uint32_t)((int32_t)((uint32_t)q.y * pc) >> 4));
                x[3] = (int32_t)((uint32_t)q.w - (uint32_t)((int32_t)((uint32_t)q.z * pc) >> 4));
                x[4] = (int32_t)((uint32_t)nxt - (uint32_t)((int32_t)((uint32_t)q.w * pc) >> 4));
            } else {
                #pragma unroll
                for (int t = 0; t < 5; ++t) {
                    const int32_t i = 4 * j + t;
                    int32_t v = 0;
                    if (i >= 0 && (uint32_t)i < n) { v = (int32_t)((uint32_t)raw[i] - (uint32_t)((int32_t)((uint32_t)raw[i ? i - 1 : 0] * pc) >> 4)); }
                    x[t] = v;
                }
            }
            #pragma unroll
            for (int t = 0; t < 4; ++t) { fits &= ((uint32_t)(x[t] + 32768) < 65536u) ? 1 : 0; }
            Z[zpair_slot(jj + zf - 1u)] = pack_pair_entry(x);
        }
    } else {
        apply_preemphasis(region_i, sig, n, pre_coef);
        __syncthreads();
        apply_ltp(sig, region_i, n, p.ltp_order, ltp_period, out->ltp_coef[0], out->ltp_coef[1], out->ltp_coef[2]);
        for (uint32_t jj = tid; jj < zcount + 1u; jj += kThreads) {
            /* sig is zero beyond n (apply_preemphasis pads 12 samples past the rounded end) and in the four samples in front */
            const int32_t j = (int32_t)jj - 1;
            const int4 q = *reinterpret_cast<const int4 *>(sig + 4 * j);
            const int32_t x[5] = { q.x, q.y, q.z, q.w, sig[4 * j + 4] };
            #pragma unroll
            for (int t = 0; t < 4; ++t) { fits &= ((uint32_t)(x[t] + 32768) < 65536u) ? 1 : 0; }
            Z[zpair_slot(jj + zf - 1u)] = pack_pair_entry(x);
        }
    }
    const bool zmode = __syncthreads_and(fits) != 0;
    if (zmode) {
        for (uint32_t m = tid; m < (p4 >> 2); m += kThreads) {
            coef_b[m] = (int32_t)(((uint32_t)coef_s[4u * m] & 0xffu) | (((uint32_t)coef_s[4u * m + 1u] & 0xffu) << 8)
                                | (((uint32_t)coef_s[4u * m + 2u] & 0xffu) << 16) | (((uint32_t)coef_s[4u * m + 3u] & 0xffu) << 24));
        }
        __syncthreads();
        fir_residual_dp2a(Z, zf, coef_b, n, order, p4, rshift, half, res_s, res_g);
    } else {
        /* ---- the signal does not fit 16 bits (24-bit sources, loud side channels): int32 signal, IMAD filter ---- */
        if (ltp_period == 0u) { apply_preemphasis(region_i, sig, n, pre_coef); }
        __syncthreads();
        fir_residual_imad(sig, coef_s, n, order, p4, rshift, half, res_s, res_g);
    }
    __syncthreads();

    residual_finish(p, out, res_s, n, scratch, red32, coef_s, p4, order, ltp_period);
}

__global__ void __launch_bounds__(kThreads, 4) residual_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    residual_item<false>(p, smem, blockIdx.x);
}

__global__ void __launch_bounds__(kThreads, 4) residual_big_kernel(const __grid_constant__ LaunchParams p)
{
    unsigned char *scratch = p.big_scratch + (size_t)blockIdx.x * p.big_stride;
    const uint32_t total = p.num_jobs * p.ncand;
    for (uint32_t item = blockIdx.x; item < total; item += gridDim.x) {
        residual_item<true>(p, scratch, item);
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------------
 * decide_kernel: one thread per job (srla_encoder.c:766-796, 1276-1327, 1477-1546, 1608-1611)
 * ---------------------------------------------------------------------------------------------- */
__global__ void decide_kernel(const __grid_constant__ LaunchParams p)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.num_jobs) { return; }
    const Job job = p.jobs[j];
    const CandOut *c = p.cand + (size_t)j * p.ncand;
    JobOut o;
    const uint32_t nch = p.nch, n = job.nsmpl, first_ch = (nch >= 2u) ? 2u : 0u;
    const uint32_t raw_bits = p.bps * n * nch;
    o.status = 0; o.method = 0; o.pad = 0; o.out_offset = 0;
    for (uint32_t ch = 0; ch < (uint32_t)kMaxChannels; ++ch) { o.cand_of_channel[ch] = first_ch + ch; }
    uint32_t type = kBlockCompress;
    if (n <= p.max_order) { type = kBlockRaw; }
    else {
        uint32_t nz = 0;
        for (uint32_t ch = 0; ch < nch; ++ch) { nz |= c[first_ch + ch].nonzero; }
        if (!nz) { type = kBlockSilent; }
    }
    uint32_t est_type = type;
    uint32_t emitted_bits = 0, payload_bits = 0;
    if (type == kBlockCompress) {
        for (uint32_t k = 0; k < p.ncand; ++k) { if (c[k].status) { o.status = 1; } }
        if (nch >= 2u) {
            const uint32_t M = c[0].total_bits, S = c[1].total_bits, Lb = c[2].total_bits, R = c[3].total_bits;
            const uint32_t cost[4] = { Lb + R, M + S, Lb + S, S + R };
            uint32_t best = 0;
            for (uint32_t m = 1; m < 4; ++m) { if (cost[best] > cost[m]) { best = m; } }
            o.method = best;
            if (best == 1u) { o.cand_of_channel[0] = 0; o.cand_of_channel[1] = 1; }
            else if (best == 2u) { o.cand_of_channel[1] = 1; }
            else if (best == 3u) { o.cand_of_channel[0] = 1; }
            payload_bits = cost[best];
        } else {
            payload_bits = c[0].total_bits;
        }
        payload_bits = (payload_bits + 2u + 7u) & ~7u;
        emitted_bits = 2u;
        for (uint32_t ch = 0; ch < nch; ++ch) { emitted_bits += c[o.cand_of_channel[ch]].total_bits; }
        emitted_bits = (emitted_bits + 7u) & ~7u;
        if (emitted_bits >= raw_bits) { type = kBlockRaw; }
        if (payload_bits >= raw_bits) { est_type = kBlockRaw; }
    }
    o.type = type;
    o.bytes = 11u + ((type == kBlockCompress) ? emitted_bits / 8u : (type == kBlockRaw) ? raw_bits / 8u : 0u);
    o.estimate_bytes = 11u + ((est_type == kBlockCompress) ? payload_bits / 8u : (est_type == kBlockRaw) ? raw_bits / 8u : 0u);
    p.jobout[j] = o;
}

/* ------------------------------------------------------------------------------------------------
 * scan_kernel: single CTA; block offsets in job order, 30-byte stream headers in front of the
 * first block of every stream.  running[0] carries the offset across launches.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(1024) scan_kernel(const __grid_constant__ LaunchParams p)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = p.running[0]; }
    __syncthreads();
    for (uint32_t base = 0; base < p.num_jobs; base += 1024u) {
        const uint32_t j = base + tid;
        unsigned long long mine = 0, hdr = 0;
        if (j < p.num_jobs) {
            hdr = (p.emit_stream_header && (p.jobs[j].flags & kJobFirstOfStream)) ? 30ull : 0ull;
            mine = (unsigned long long)p.jobout[j].bytes + hdr;
        }
        unsigned long long x = mine;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) { x += y; } }
        if (lane == 31) { warp_tot[warp] = x; }
        __syncthreads();
        unsigned long long pre = carry;
        for (int w = 0; w < warp; ++w) { pre += warp_tot[w]; }
        if (j < p.num_jobs) {
            const unsigned long long start = pre + x - mine;     /* where the (optional) stream header begins */
            p.jobout[j].out_offset = start + hdr;
            if (p.jobs[j].flags & kJobFirstOfStream) { p.stream_begin[p.jobs[j].stream] = start; }
        }
        __syncthreads();
        if (tid == 1023) { carry = pre + x; }
        __syncthreads();
    }
    if (tid == 0) {
        p.running[0] = carry;
        if (carry > p.out_capacity) { p.running[1] = 1ull; }
    }
}

/* ------------------------------------------------------------------------------------------------
 * emit_kernel: one CTA per job.  The block is assembled in shared memory as big-endian 32-bit
 * words (bit 31 of word 0 is the first bit of the block), then copied out.
 * ---------------------------------------------------------------------------------------------- */
struct BitCursor {
    uint32_t *words;          /* staging */
    uint32_t lo_excl, hi_excl;/* words in [lo_excl, hi_excl) are owned exclusively by this thread */
    uint32_t cur;             /* word index being accumulated */
    uint32_t acc;             /* bits of that word */
    __device__ __forceinline__ void init(uint32_t *w, uint32_t start_bit, uint32_t end_bit)
    {
        words = w; lo_excl = (start_bit + 31u) >> 5; hi_excl = end_bit >> 5; cur = start_bit >> 5; acc = 0;
    }
    __device__ __forceinline__ void flush()
    {
        if (acc) {
            if (cur >= lo_excl && cur < hi_excl) { words[cur] = acc; } else { atomicOr(words + cur, acc); }
        }
    }
    __device__ __forceinline__ void seek_word(uint32_t w) { if (w != cur) { flush(); cur = w; acc = 0; } }
    /* put the low nbits of value (nbits <= 32) at absolute bit position pos */
    __device__ __forceinline__ void put(uint32_t pos, uint32_t value, uint32_t nbits)
    {
        if (nbits == 0u) { return; }
        if (nbits < 32u) { value &= (1u << nbits) - 1u; }
        const uint32_t w = pos >> 5, off = pos & 31u;
        seek_word(w);
        const uint32_t room = 32u - off;
        if (nbits <= room) { acc |= value << (room - nbits); }
        else {
            acc |= value >> (nbits - room);
            seek_word(w + 1u);
            acc |= value << (32u - (nbits - room));
        }
    }
};

/* sequential MSB-first writer of one thread's contiguous bit range (bit_stream.h:245-307 semantics): a 64-bit
 * shift register whose completed upper word is stored as soon as it fills.  Only the first and the last word
 * of the range can be shared with the neighbouring threads: those are OR-ed atomically, every other word is
 * owned exclusively and stored plainly (the staging buffer starts zeroed). */
struct BitSink {
    uint32_t *words;
    unsigned long long acc;
    uint32_t w, first_w, fill;
    __device__ __forceinline__ void init(uint32_t *base, uint32_t start_bit) { words = base; w = start_bit >> 5; first_w = w; fill = start_bit & 31u; acc = 0ull; }
    __device__ __forceinline__ void flush_word()
    {
        const uint32_t out = (uint32_t)(acc >> 32);
        if (w == first_w) { if (out) { atomicOr(words + w, out); } } else { words[w] = out; }
        acc <<= 32; fill -= 32u; w++;
    }
    __device__ __forceinline__ void zeros(uint32_t q) { fill += q; while (fill >= 32u) { flush_word(); } }
    /* the low len bits of value, 1 <= len <= 32, value < 2^len */
    __device__ __forceinline__ void put(uint32_t value, uint32_t len)
    {
        acc |= (unsigned long long)value << (64u - fill - len);
        fill += len;
        if (fill >= 32u) { flush_word(); }
    }
    __device__ __forceinline__ void finish() { const uint32_t out = (uint32_t)(acc >> 32); if (out) { atomicOr(words + w, out); } }
};

/* length in bits of the code of u with parameter k (srla_coder.c:165-190) */
__device__ __forceinline__ uint32_t code_len(uint32_t u, uint32_t k, uint32_t code_type)
{
    if (code_type == kCodeRice) { return 1u + k + (u >> k); }
    const uint32_t k1 = k + 1u, pivot = 1u << k1;
    return (u < pivot) ? (k1 + 1u) : (2u + ((u - pivot) >> k) + k);
}
__device__ __forceinline__ uint32_t emit_code(BitCursor &bc, uint32_t pos, uint32_t u, uint32_t k, uint32_t code_type)
{
    if (code_type == kCodeRice) {
        const uint32_t q = u >> k;
        bc.put(pos + q, (1u << k) | (u & ((1u << k) - 1u)), k + 1u);
        return q + 1u + k;
    }
    const uint32_t k1 = k + 1u, pivot = 1u << k1;
    if (u < pivot) { bc.put(pos, pivot | u, k1 + 1u); return k1 + 1u; }
    const uint32_t z = u - pivot, q = 1u + (z >> k);
    bc.put(pos + q, (1u << k) | (z & ((1u << k) - 1u)), k + 1u);
    return q + 1u + k;
}

/* one block; `words`: the staging area (shared memory, or -- blocks beyond its capacity -- a per-CTA area in global memory) */
__device__ __forceinline__ void emit_block(const LaunchParams &p, uint32_t *words, const uint32_t j, uint32_t *scan_scratch, uint32_t *fl_lo, uint32_t *fl_hi)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Job job = p.jobs[j];
    const JobOut jo = p.jobout[j];
    const StreamDev st = p.streams[job.stream];
    const uint32_t n = job.nsmpl, nch = p.nch, bps = p.bps;
    const uint32_t lshift = p.use_fixed_lshift ? p.fixed_lshift : st.lshift;
    const uint32_t nbytes = jo.bytes;
    const bool fits = (jo.out_offset + nbytes <= p.out_capacity) && (nbytes <= p.emit_smem_bytes) && (jo.status == 0u);

    /* stream header (srla_encoder.c:85-165) */
    if (p.emit_stream_header && (job.flags & kJobFirstOfStream) && tid == 0 && jo.out_offset >= 30ull && jo.out_offset <= p.out_capacity) {
        uint8_t *h = p.out + (jo.out_offset - 30ull);
        const uint32_t ns = st.num_samples;
        h[0] = '1'; h[1] = '2'; h[2] = '4'; h[3] = '9';
        h[4] = 0; h[5] = 0; h[6] = 0; h[7] = 10;
        h[8] = 0; h[9] = 0; h[10] = 0; h[11] = 18;
        h[12] = (uint8_t)(nch >> 8); h[13] = (uint8_t)nch;
        h[14] = (uint8_t)(ns >> 24); h[15] = (uint8_t)(ns >> 16); h[16] = (uint8_t)(ns >> 8); h[17] = (uint8_t)ns;
        h[18] = (uint8_t)(p.sampling_rate >> 24); h[19] = (uint8_t)(p.sampling_rate >> 16); h[20] = (uint8_t)(p.sampling_rate >> 8); h[21] = (uint8_t)p.sampling_rate;
        h[22] = (uint8_t)(bps >> 8); h[23] = (uint8_t)bps;
        h[24] = (uint8_t)lshift;
        h[25] = (uint8_t)(p.max_block >> 24); h[26] = (uint8_t)(p.max_block >> 16); h[27] = (uint8_t)(p.max_block >> 8); h[28] = (uint8_t)p.max_block;
        h[29] = (uint8_t)p.preset;
    }
    if (!fits) { return; }

    const uint32_t nwords = (nbytes + 3u) >> 2;
    for (uint32_t w = tid; w < nwords + 1u; w += kThreads) { words[w] = 0u; }
    __syncthreads();

    if (jo.type == kBlockRaw) {
        /* interleaved big-endian zig-zag samples of the UNSHIFTED input (srla_encoder.c:799-858) */
        const uint32_t bytes_per = bps >> 3;
        for (uint32_t e = tid; e < n * nch; e += kThreads) {
            const uint32_t i = e / nch, ch = e % nch;
            const uint32_t v = zigzag32(load_sample(st, ch, job.offset + i));
            const uint32_t bytepos = 11u + e * bytes_per;
            for (uint32_t b = 0; b < bytes_per; ++b) {
                const uint32_t byte = (v >> (8u * (bytes_per - 1u - b))) & 0xffu;
                const uint32_t at = bytepos + b;
                atomicOr(words + (at >> 2), byte << (24u - 8u * (at & 3u)));
            }
        }
    } else if (jo.type == kBlockCompress) {
        const CandOut *cands = p.cand + (size_t)j * p.ncand;
        const uint32_t base_bit = 88u;
        /* section 1: method, pre-emphasis state (srla_encoder.c:1369-1386) */
        if (tid == 0) {
            BitCursor bc; bc.init(words, 0u, 0u);
            uint32_t pos = base_bit;
            bc.put(pos, jo.method, 2u); pos += 2u;
            for (uint32_t ch = 0; ch < nch; ++ch) {
                const CandOut &c = cands[jo.cand_of_channel[ch]];
                /* bps + 1 may be 33 bits only for bps 32 which the format's raw blocks exclude */
                bc.put(pos, zigzag32(c.pre_prev), bps + 1u); pos += bps + 1u;
                bc.put(pos, zigzag32(c.pre_coef), 5u); pos += 5u;
            }
            bc.flush();
        }
        uint32_t pos = base_bit + 2u + nch * (bps + 1u + 5u);
        /* section 2: LPC parameters, Huffman-coded coefficients (srla_encoder.c:1388-1419) */
        for (uint32_t ch = 0; ch < nch; ++ch) {
            const CandOut &c = cands[jo.cand_of_channel[ch]];
            const uint32_t order = c.order;
            if (tid == 0) {
                BitCursor bc; bc.init(words, 0u, 0u);
                bc.put(pos, order, 8u); bc.put(pos + 8u, c.rshift, 4u); bc.put(pos + 12u, c.use_sum, 1u);
                bc.flush();
            }
            uint32_t code = 0, len = 0;
            if ((uint32_t)tid < order) {
                const int32_t cf = c.coef[tid];
                uint32_t sym, table = 0;
                if (tid == 0 || !c.use_sum) { sym = zigzag32(cf); } else { sym = zigzag32(cf + (int32_t)c.coef[tid - 1]); table = 256u; }
                code = __ldg(p.huff_code + table + sym); len = __ldg(p.huff_len + table + sym);
            }
            uint32_t total;
            const uint32_t incl = block_scan_inclusive(len, scan_scratch, &total);
            if (len) { BitCursor bc; bc.init(words, 0u, 0u); bc.put(pos + 13u + incl - len, code, len); bc.flush(); }
            pos += 13u + total;
        }
        /* section 3: LTP parameters (srla_encoder.c:1421-1438) */
        if (tid == 0) {
            BitCursor bc; bc.init(words, 0u, 0u);
            uint32_t q = pos;
            for (uint32_t ch = 0; ch < nch; ++ch) {
                const CandOut &c = cands[jo.cand_of_channel[ch]];
                bc.put(q, c.ltp_period != 0u, 1u); q += 1u;
                if (c.ltp_period) {
                    bc.put(q, (p.ltp_order - 1u) / 2u, 1u); q += 1u;
                    bc.put(q, c.ltp_period - (uint32_t)kLtpMinPeriod, 8u); q += 8u;
                    for (uint32_t t = 0; t < p.ltp_order; ++t) { bc.put(q, zigzag32(c.ltp_coef[t]), 6u); q += 6u; }
                }
            }
            bc.flush();
        }
        for (uint32_t ch = 0; ch < nch; ++ch) { const CandOut &c = cands[jo.cand_of_channel[ch]]; pos += 1u + (c.ltp_period ? (1u + 8u + 6u * p.ltp_order) : 0u); }
        __syncthreads();
        /* section 4: residual codes (srla_coder.c:486-595).  Every thread codes one contiguous run of samples,
         * four at a time (one 16-byte load): pass 1 sizes the run, a block scan places it, pass 2 writes it.
         * The loops over the quads are deliberately not unrolled: the body is large and the kernel is otherwise
         * bound by instruction fetch. */
        const uint32_t chunk = (n + kThreads - 1u) / kThreads;
        for (uint32_t ch = 0; ch < nch; ++ch) {
            const uint32_t cidx = jo.cand_of_channel[ch];
            const CandOut &c = cands[cidx];
            const uint32_t code_type = c.code_type;
            if (tid == 0) {
                BitCursor bc; bc.init(words, 0u, 0u);
                bc.put(pos, code_type, 2u);
                if (code_type != kCodeAllZero) { bc.put(pos + 2u, c.porder, 10u); }
                bc.flush();
            }
            if (code_type == kCodeAllZero) { pos += 2u; __syncthreads(); continue; }
            const int32_t *res = p.residual + ((size_t)j * p.ncand + cidx) * p.res_stride;
            const uint32_t plen = n >> c.porder;
            const uint32_t i0 = (uint32_t)tid * chunk, i1 = (i0 + chunk < n) ? i0 + chunk : n;
            const uint32_t cnt = (i0 < n) ? i1 - i0 : 0u;
            const bool vec = (chunk & 3u) == 0u;                 /* runs start 16-byte aligned; res_stride pads the last quad */
            /* partitions: `part0` holds sample i0; a partition's parameter field precedes its first sample */
            const uint32_t part0 = (cnt > 0u) ? i0 / plen : 0u;
            const uint32_t first_boundary = (part0 * plen == i0) ? i0 : (part0 + 1u) * plen;
            const bool rice = (code_type == kCodeRice);
            /* pass 1: the residual stage left this thread's bit count when its sample ranges are the same as here */
            uint32_t mybits = 0;
            const uint32_t known = (c.tb_valid && (n & (kThreads - 1u)) == 0u) ? (uint32_t)c.thread_bits[tid] : 0xffffu;
            if (known != 0xffffu) { mybits = known; }
            else {
                uint32_t part = part0, k = (cnt > 0u) ? (uint32_t)c.kparam[part0] : 0u, next = first_boundary;
                #pragma unroll 1
                for (uint32_t t0 = 0; t0 < cnt; t0 += 4u) {
                    int4 q4 = make_int4(0, 0, 0, 0);
                    if (vec) { q4 = __ldg(reinterpret_cast<const int4 *>(res + i0 + t0)); }
                    else {
                        q4.x = __ldg(res + i0 + t0);
                        if (t0 + 1u < cnt) { q4.y = __ldg(res + i0 + t0 + 1u); }
                        if (t0 + 2u < cnt) { q4.z = __ldg(res + i0 + t0 + 2u); }
                        if (t0 + 3u < cnt) { q4.w = __ldg(res + i0 + t0 + 3u); }
                    }
                    const uint32_t uu[4] = { zigzag32(q4.x), zigzag32(q4.y), zigzag32(q4.z), zigzag32(q4.w) };
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t i = i0 + t0 + (uint32_t)e;
                        if (t0 + (uint32_t)e < cnt) {
                            if (i == next) {
                                part = (i == i0) ? part0 : part + 1u;
                                k = c.kparam[part];
                                mybits += (part == 0u) ? 5u : (zigzag32((int32_t)k - (int32_t)c.kparam[part - 1u]) + 1u);
                                next += plen;
                            }
                            mybits += code_len(uu[e], k, code_type);
                        }
                    }
                }
            }
            uint32_t total;
            const uint32_t incl = block_scan_inclusive(mybits, scan_scratch, &total);
            /* pass 2 */
            if (mybits) {
                BitSink bs; bs.init(words, pos + 12u + incl - mybits);
                uint32_t part = part0, k = c.kparam[part0], next = first_boundary;
                #pragma unroll 1
                for (uint32_t t0 = 0; t0 < cnt; t0 += 4u) {
                    int4 q4 = make_int4(0, 0, 0, 0);
                    if (vec) { q4 = __ldg(reinterpret_cast<const int4 *>(res + i0 + t0)); }
                    else {
                        q4.x = __ldg(res + i0 + t0);
                        if (t0 + 1u < cnt) { q4.y = __ldg(res + i0 + t0 + 1u); }
                        if (t0 + 2u < cnt) { q4.z = __ldg(res + i0 + t0 + 2u); }
                        if (t0 + 3u < cnt) { q4.w = __ldg(res + i0 + t0 + 3u); }
                    }
                    const uint32_t uu[4] = { zigzag32(q4.x), zigzag32(q4.y), zigzag32(q4.z), zigzag32(q4.w) };
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t i = i0 + t0 + (uint32_t)e;
                        if (t0 + (uint32_t)e < cnt) {
                            if (i == next) {
                                part = (i == i0) ? part0 : part + 1u;
                                k = c.kparam[part];
                                if (part == 0u) { bs.put(k, 5u); }
                                else { bs.zeros(zigzag32((int32_t)k - (int32_t)c.kparam[part - 1u])); bs.put(1u, 1u); }
                                next += plen;
                            }
                            /* q zeros, a one, nb low bits (srla_coder.c:165-190) */
                            const uint32_t v = uu[e];
                            uint32_t q, nb, low;
                            if (rice) { q = v >> k; nb = k; low = v & ((1u << k) - 1u); }
                            else {
                                const uint32_t k1 = k + 1u, pivot = 1u << k1;
                                const bool small = v < pivot;
                                const uint32_t z = v - pivot;
                                q = small ? 0u : 1u + (z >> k);
                                nb = small ? k1 : k;
                                low = small ? v : (z & ((1u << k) - 1u));
                            }
                            if (q + nb < 32u) { bs.put((1u << nb) | low, q + nb + 1u); }      /* leading zeros ride along */
                            else {
                                bs.zeros(q);
                                if (nb < 32u) { bs.put((1u << nb) | low, nb + 1u); } else { bs.put(1u, 1u); bs.put(low, 32u); }
                            }
                        }
                    }
                }
                bs.finish();
            }
            pos += 12u + total;
            __syncthreads();
        }
    }
    __syncthreads();

    /* block header (srla_encoder.c:1585-1595, 1629-1636): sync, size, checksum, type, nsmpl */
    if (tid == 0) {
        const uint32_t size_field = nbytes - 11u + 5u;
        words[0] = 0xFFFF0000u | (size_field >> 16);
        words[1] = (size_field << 16);                       /* checksum patched below */
        words[2] |= (jo.type << 24) | ((n & 0xffffu) << 8);
    }
    __syncthreads();
    /* Fletcher-16 over bytes [8, nbytes) (srla_utility.c:36-60): lo = sum d, hi = sum (L - i) d, mod 255.
     * Bytes 8.. start on a word boundary of the staging buffer; partial sums stay below 2^32 for 64 words
     * (4 * 64 * 255 * 2^16), so the modulo is taken once per 32 words. */
    {
        const uint32_t Lb = nbytes - 8u;
        const uint32_t nw = (Lb + 3u) >> 2;
        uint32_t lo = 0, hi = 0, since = 0;
        for (uint32_t wi = tid; wi < nw; wi += kThreads) {
            const uint32_t wd = words[2u + wi];               /* bytes beyond nbytes are zero */
            const uint32_t i = wi << 2;                        /* index of the word's first byte inside [8, nbytes) */
            const uint32_t d0 = wd >> 24, d1 = (wd >> 16) & 0xffu, d2 = (wd >> 8) & 0xffu, d3 = wd & 0xffu;
            lo += d0 + d1 + d2 + d3;
            /* weights (Lb - i - t); a zero padding byte may get a "negative" weight, times zero */
            const uint32_t wgt = Lb - i;
            hi += wgt * d0 + (wgt - 1u) * d1 + (wgt - 2u) * d2 + (wgt - 3u) * d3;
            if (++since == 32u) { hi %= 255u; since = 0; }
        }
        lo %= 255u; hi %= 255u;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, o); hi += __shfl_xor_sync(0xffffffffu, hi, o); }
        if (lane == 0) { fl_lo[warp] = lo % 255u; fl_hi[warp] = hi % 255u; }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = 0, b = 0;
            for (int w = 0; w < kWarps; ++w) { a += fl_lo[w]; b += fl_hi[w]; }
            a %= 255u; b %= 255u;
            words[1] |= (b << 8) | a;
        }
        __syncthreads();
    }
    /* copy out: byte b of the block = big-endian byte b of the staging words */
    {
        uint8_t *dst = p.out + jo.out_offset;
        const uint32_t mis = (uint32_t)((4u - (reinterpret_cast<unsigned long long>(dst) & 3ull)) & 3ull);
        const uint32_t head = (mis < nbytes) ? mis : nbytes;
        const uint32_t body_words = (nbytes - head) >> 2;
        const uint32_t tail_at = head + (body_words << 2);
        if ((uint32_t)tid < head) { dst[tid] = (uint8_t)(words[tid >> 2] >> (24u - 8u * (tid & 3u))); }
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(dst + head);
        for (uint32_t w = tid; w < body_words; w += kThreads) {
            const uint32_t s = head + (w << 2);
            const uint32_t be = __funnelshift_l(words[(s >> 2) + 1u], words[s >> 2], 8u * (s & 3u));
            dst32[w] = __byte_perm(be, 0u, 0x0123);
        }
        if ((uint32_t)tid < nbytes - tail_at) { const uint32_t at = tail_at + tid; dst[at] = (uint8_t)(words[at >> 2] >> (24u - 8u * (at & 3u))); }
    }
    /* statistics */
    if (p.stats && tid == 0) {
        atomicAdd(p.stats + 256 + 4 + jo.type, 1u);
        if (jo.type == kBlockCompress) {
            atomicAdd(p.stats + 256 + jo.method, 1u);
            const CandOut *cands = p.cand + (size_t)j * p.ncand;
            for (uint32_t ch = 0; ch < nch; ++ch) { atomicAdd(p.stats + (cands[jo.cand_of_channel[ch]].order & 255u), 1u); }
        }
    }
}

__global__ void __launch_bounds__(kThreads) emit_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t scan_scratch[kWarps + 1];
    __shared__ uint32_t fl_lo[kWarps], fl_hi[kWarps];
    emit_block(p, reinterpret_cast<uint32_t *>(smem), blockIdx.x, scan_scratch, fl_lo, fl_hi);
}

/* blocks whose encoded size exceeds the shared memory of an SM: persistent CTAs, staging area in global memory */
__global__ void __launch_bounds__(kThreads) emit_big_kernel(const __grid_constant__ LaunchParams p)
{
    __shared__ uint32_t scan_scratch[kWarps + 1];
    __shared__ uint32_t fl_lo[kWarps], fl_hi[kWarps];
    uint32_t *words = reinterpret_cast<uint32_t *>(p.big_scratch + (size_t)blockIdx.x * p.big_stride);
    for (uint32_t j = blockIdx.x; j < p.num_jobs; j += gridDim.x) {
        emit_block(p, words, j, scan_scratch, fl_lo, fl_hi);
        __syncthreads();
    }
}

} // namespace srla
#endif
