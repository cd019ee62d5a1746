/*
 * libsrla_b200.cu -- host runtime and C ABI of the B200-native SRLA encode path.
 *
 * Exports the encoder half of the reference's public API (include/srla_encoder.h) plus the batch
 * extension declared in include/srla_b200.h.  All signal processing runs in the CUDA kernels of
 * kernels.cuh; the host side validates arguments exactly like the reference, builds the job list
 * (one job per block), launches analyse -> decide -> scan -> emit per batch, and -- in variable-block
 * mode -- runs the reference's shortest-path block division (srla_encoder.c:249-424) on the exact
 * candidate sizes the GPU computed.
 *
 * There is deliberately no CPU encode path in this file: without a CUDA device Create() fails.
 */
#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <map>
#include <new>
#include <vector>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

#if defined(__linux__)
#include <sched.h>
#endif

#include <cuda_runtime.h>

#include "../../include/srla_b200.h"
#include "host_tables.h"
#include "kernels.cuh"

using namespace srla;

namespace {

const uint32_t kPresetMaxOrder[SRLA_NUM_PARAMETER_PRESETS] = { 0, 8, 16, 32, 64, 128, 255 };   /* srla_internal.c:30-38 */
const uint32_t kEncoderMagic = 0x53424C41u;
thread_local int g_device = -1;

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            std::fprintf(stderr, "[srla_b200] %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(e_), \
                         __FILE__, __LINE__);                                                          \
            return false;                                                                              \
        }                                                                                              \
    } while (0)

/* growable device / pinned-host buffers */
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    bool reserve(size_t bytes)
    {
        if (bytes <= cap) { return true; }
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { std::fprintf(stderr, "[srla_b200] cudaMalloc(%zu) failed\n", want); p = nullptr; return false; }
        cap = want;
        return true;
    }
    void release() { if (p) { cudaFree(p); } p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr; size_t cap = 0;
    bool reserve(size_t bytes)
    {
        if (bytes <= cap) { return true; }
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) { std::fprintf(stderr, "[srla_b200] cudaMallocHost(%zu) failed\n", want); p = nullptr; return false; }
        cap = want;
        return true;
    }
    void release() { if (p) { cudaFreeHost(p); } p = nullptr; cap = 0; }
};

/* Host threads of a handle, created once and parked between calls: spawning sixteen threads per EncodeWhole call
 * cost ~2 ms on the measured box, a quarter of the call. */
struct WorkerPool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable wake, idle;
    std::function<void()> job;
    unsigned long long epoch = 0;
    int running = 0;
    bool quit = false;
    void ensure(int n)
    {
        while ((int)threads.size() < n) { threads.emplace_back([this] { loop(); }); }
    }
    void loop()
    {
        unsigned long long seen = 0;
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m);
                wake.wait(l, [&] { return quit || epoch != seen; });
                if (quit) { return; }
                seen = epoch; f = job;
            }
            f();
            { std::lock_guard<std::mutex> l(m); if (--running == 0) { idle.notify_all(); } }
        }
    }
    void start(std::function<void()> f)                       /* every thread runs f once */
    {
        { std::lock_guard<std::mutex> l(m); job = std::move(f); running = (int)threads.size(); ++epoch; }
        wake.notify_all();
    }
    void wait() { std::unique_lock<std::mutex> l(m); idle.wait(l, [&] { return running == 0; }); }
    ~WorkerPool()
    {
        { std::lock_guard<std::mutex> l(m); quit = true; }
        wake.notify_all();
        for (std::thread &t : threads) { if (t.joinable()) { t.join(); } }
    }
};

/* cudaFuncSetAttribute state per (device, kernel), process-wide (see Runner::prep_kernel) */
struct FuncAttr { uint32_t max_dynamic = 0; int carveout = -2; };
struct FuncAttrCache { std::mutex m; std::map<std::pair<int, const void *>, FuncAttr> set; };
FuncAttrCache g_func_attr;

constexpr int kMaxLanes = 4;
constexpr size_t kMiscBytes = 2 * sizeof(unsigned long long) + 263 * sizeof(uint32_t) + 4;   /* running[2], stats[263], pad to 8 */

struct DeviceCtx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_upload = nullptr;
    bool upload_pending = false;
    std::vector<cudaEvent_t> ev_pool;      /* per batch: front begin, lpc begin, residual begin, decide begin, end */
    size_t ev_used = 0;
    /* tables */
    DevBuf tw_complex, tw_real, rice_thr, huff_code, huff_len;
    uint32_t tw_c_off[20], tw_r_off[20];
    /* work */
    DevBuf streams, jobs, misc, stream_begin, pcm, out, raw;
    /* a lane = one compute stream + its own per-group scratch; alternate groups of a call run on different
     * lanes so the latency-bound kernels of one group (lpc, scan) overlap the throughput-bound ones of the next */
    struct Lane { cudaStream_t own = nullptr, stream = nullptr; cudaEvent_t done = nullptr; DevBuf cand, diag, jobout, residual, lags, lpc_state, svr_coef, svr_matrix, big; };
    Lane lane[kMaxLanes];
    int lanes = 3;                 /* SRLA_B200_LANES */
    int groups = 8;                /* SRLA_B200_GROUPS: groups a large call is split into */
    int sized_carveout = 1;        /* SRLA_B200_CARVE=max: every kernel asks for the maximum shared-memory carve-out */
    int ramp = 1;                  /* SRLA_B200_RAMP=0: uniform groups on the host path */
    int trace = 0;                 /* SRLA_B200_TRACE=1: print the per-group timeline of every pipelined call */
    int split_device = 3;          /* SRLA_B200_SPLIT_DEVICE=n: a large device-resident call runs as n groups on n lanes (0 / 1: one batch on the
                                      caller's stream).  The FP64-bound front kernel of one group then overlaps the lpc / residual / emit
                                      kernels of another: 2.60 -> 2.47 ms per config-2 step with 2, 3 or 4 groups alike */
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_scan;
    PinBuf h_jobs, h_small, h_jobout, h_result, h_mailbox, h_stage, h_stage_out;
    int feed_threads = 8;          /* SRLA_B200_FEED_THREADS */
    std::unique_ptr<WorkerPool> pool;
    DevBuf snapshot;
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_h2d, ev_grp, ev_d2h;
    std::vector<Job> jobs_scratch;
    std::vector<TailJob> tails_scratch;   /* tail blocks of the fixed tiling that need front_tail_kernel (odd, or short with LTP) */
    DevBuf tails, tail_scratch, chain;   /* chain: variable blocks -- the reference's call list around a stream's end (front_tail_kernel) */
    std::vector<uint32_t> jobs_key;   /* stream lengths + block size of the fixed tiling that jobs_scratch and the device copy hold */
    bool jobs_cached = false;
    int max_smem_optin = 0;
    int num_sms = 0;
    int resid16 = 1;               /* SRLA_B200_RESID16=0: the one-CTA-per-candidate residual kernel for 16-bit PCM as well (tuning / A-B) */
    int front_occ = 3;             /* CTAs per SM front_kernel<128> is register-sized for (SRLA_B200_FRONT_OCC=3|4, tuning) */
    int front16 = 1;               /* SRLA_B200_FRONT16=0: front_kernel for 16-bit PCM as well (tuning / A-B) */
};

} // namespace

struct SRLAEncoder {
    uint32_t magic;
    struct SRLAEncoderConfig config;
    struct SRLAEncodeParameter param;
    int set_parameter;
    uint32_t max_order;
    uint8_t offset_lshift;           /* header.offset_lshift of the reference handle */
    uint8_t alloced_by_own;
    void *work;
    DeviceCtx *ctx;
    struct SRLAB200Stats stats;
};

namespace {

/* CPUs this process may run on (its affinity mask, not the machine's CPU count): a rank pinned next to its GPU sizes its
 * feeder team for its own share of the host */
unsigned usable_cpus()
{
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int n = CPU_COUNT(&set); if (n > 0) { return (unsigned)n; } }
#endif
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? hc : 8u;
}

/* the library carries sm_100a code only (arch-specific: it does not run on any other compute capability) */
bool device_is_sm100(int device)
{
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess
        || cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) { return false; }
    if (major != 10 || minor != 0) {
        std::fprintf(stderr, "[srla_b200] device %d is sm_%d%d; this library is built for sm_100a only\n", device, major, minor);
        return false;
    }
    return true;
}

bool ctx_init(DeviceCtx *c)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        std::fprintf(stderr, "[srla_b200] no CUDA device: the SRLA B200 encode path has no CPU fallback\n");
        return false;
    }
    if (g_device >= 0) { CU_TRY(cudaSetDevice(g_device)); }
    CU_TRY(cudaGetDevice(&c->device));
    if (!device_is_sm100(c->device)) { return false; }
    CU_TRY(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
    if (const char *e = std::getenv("SRLA_B200_FRONT_OCC")) { if (e[0] >= '2' && e[0] <= '4') { c->front_occ = e[0] - '0'; } }
    CU_TRY(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    CU_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    c->lane[0].stream = c->own_stream;
    for (int l = 1; l < kMaxLanes; l++) { CU_TRY(cudaStreamCreateWithFlags(&c->lane[l].own, cudaStreamNonBlocking)); c->lane[l].stream = c->lane[l].own; }
    for (int l = 0; l < kMaxLanes; l++) { CU_TRY(cudaEventCreateWithFlags(&c->lane[l].done, cudaEventDisableTiming)); }
    CU_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    if (const char *e = std::getenv("SRLA_B200_LANES")) { const int v = std::atoi(e); if (v >= 1 && v <= kMaxLanes) { c->lanes = v; } }
    if (const char *e = std::getenv("SRLA_B200_CARVE")) { if (e[0] == 'm') { c->sized_carveout = 0; } }
    if (const char *e = std::getenv("SRLA_B200_RAMP")) { c->ramp = std::atoi(e); }
    { const unsigned cpus = usable_cpus(); c->feed_threads = (int)std::max(2u, std::min(16u, cpus >= 8u ? cpus - 2u : cpus)); }   /* two CPUs stay free for the caller's thread and the driver's: a full team sometimes loses a 3 ms time slice (measured) */
    if (const char *e = std::getenv("SRLA_B200_FEED_THREADS")) { const int v = std::atoi(e); if (v >= 0 && v <= 64) { c->feed_threads = v; } }
    if (const char *e = std::getenv("SRLA_B200_TRACE")) { c->trace = std::atoi(e); }
    if (const char *e = std::getenv("SRLA_B200_RESID16")) { c->resid16 = std::atoi(e); }
    if (const char *e = std::getenv("SRLA_B200_FRONT16")) { c->front16 = std::atoi(e); }
    if (const char *e = std::getenv("SRLA_B200_SPLIT_DEVICE")) { const int v = std::atoi(e); if (v >= 0 && v <= kMaxLanes) { c->split_device = v; } }
    if (const char *e = std::getenv("SRLA_B200_GROUPS")) { const int v = std::atoi(e); if (v >= 1 && v <= 64) { c->groups = v; } }
    CU_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreate(&c->ev_begin));
    CU_TRY(cudaEventCreate(&c->ev_end));
    CU_TRY(cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming));

    /* host-libm tables (host_tables.h) */
    std::vector<host::Cx> tab; std::vector<uint32_t> off;
    const int max_lg = 16;                                    /* real transforms up to 65536 points (blocks up to 65535 samples) */
    host::build_complex_twiddles(max_lg, tab, off);
    if (!c->tw_complex.reserve(tab.size() * sizeof(host::Cx))) { return false; }
    CU_TRY(cudaMemcpy(c->tw_complex.p, tab.data(), tab.size() * sizeof(host::Cx), cudaMemcpyHostToDevice));
    for (int i = 0; i < 20; i++) { c->tw_c_off[i] = off[i]; }
    if (!host::build_real_twiddles(max_lg, tab, off)) {
        std::fprintf(stderr, "[srla_b200] host libm: inverse real-FFT twiddles are not the conjugate of the forward ones\n");
        return false;
    }
    if (!c->tw_real.reserve(tab.size() * sizeof(host::Cx))) { return false; }
    CU_TRY(cudaMemcpy(c->tw_real.p, tab.data(), tab.size() * sizeof(host::Cx), cudaMemcpyHostToDevice));
    for (int i = 0; i < 20; i++) { c->tw_r_off[i] = off[i]; }
    double thr[32];
    host::build_rice_thresholds(thr);
    if (!c->rice_thr.reserve(sizeof(thr))) { return false; }
    CU_TRY(cudaMemcpy(c->rice_thr.p, thr, sizeof(thr), cudaMemcpyHostToDevice));
    host::HuffTable plain, summed;
    host::build_format_huffman(plain, summed);
    uint32_t codes[512]; uint8_t lens[512];
    for (int i = 0; i < 256; i++) { codes[i] = plain.code[i]; lens[i] = plain.len[i]; codes[256 + i] = summed.code[i]; lens[256 + i] = summed.len[i]; }
    if (!c->huff_code.reserve(sizeof(codes)) || !c->huff_len.reserve(sizeof(lens))) { return false; }
    CU_TRY(cudaMemcpy(c->huff_code.p, codes, sizeof(codes), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(c->huff_len.p, lens, sizeof(lens), cudaMemcpyHostToDevice));
    if (!c->misc.reserve(2 * sizeof(unsigned long long) + 263 * sizeof(uint32_t) + 64)) { return false; }
    return true;
}

void ctx_destroy(DeviceCtx *c)
{
    if (!c) { return; }
    cudaSetDevice(c->device);
    if (c->own_stream) { cudaStreamSynchronize(c->own_stream); }
    for (cudaEvent_t e : c->ev_pool) { cudaEventDestroy(e); }
    for (cudaEvent_t e : c->ev_h2d) { cudaEventDestroy(e); }
    for (cudaEvent_t e : c->ev_grp) { cudaEventDestroy(e); }
    for (cudaEvent_t e : c->ev_d2h) { cudaEventDestroy(e); }
    for (cudaEvent_t e : c->ev_scan) { cudaEventDestroy(e); }
    for (int l = 0; l < kMaxLanes; l++) {
        if (c->lane[l].own) { cudaStreamSynchronize(c->lane[l].own); cudaStreamDestroy(c->lane[l].own); }
        if (c->lane[l].done) { cudaEventDestroy(c->lane[l].done); }
        DevBuf *lb[] = { &c->lane[l].cand, &c->lane[l].diag, &c->lane[l].jobout, &c->lane[l].residual, &c->lane[l].lags, &c->lane[l].lpc_state, &c->lane[l].svr_coef, &c->lane[l].svr_matrix, &c->lane[l].big };
        for (DevBuf *b : lb) { b->release(); }
    }
    if (c->ev_fork) { cudaEventDestroy(c->ev_fork); }
    if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); }
    if (c->d2h_stream) { cudaStreamDestroy(c->d2h_stream); }
    c->snapshot.release(); c->h_mailbox.release();
    if (c->ev_begin) { cudaEventDestroy(c->ev_begin); }
    if (c->ev_end) { cudaEventDestroy(c->ev_end); }
    if (c->ev_upload) { cudaEventDestroy(c->ev_upload); }
    DevBuf *bufs[] = { &c->chain, &c->tails, &c->tail_scratch, &c->tw_complex, &c->tw_real, &c->rice_thr, &c->huff_code, &c->huff_len, &c->streams, &c->jobs,
                       &c->misc, &c->stream_begin, &c->pcm, &c->out, &c->raw };
    for (DevBuf *b : bufs) { b->release(); }
    c->h_jobs.release(); c->h_small.release(); c->h_jobout.release(); c->h_result.release(); c->h_stage.release(); c->h_stage_out.release();
    if (c->own_stream) { cudaStreamDestroy(c->own_stream); }
}

bool config_valid(const struct SRLAEncoderConfig *cfg)
{
    if (!cfg) { return false; }
    if (cfg->max_num_samples_per_block == 0 || cfg->min_num_samples_per_block == 0
        || cfg->max_num_lookahead_samples == 0 || cfg->max_num_channels == 0) { return false; }
    if (cfg->max_num_parameters > cfg->max_num_samples_per_block) { return false; }
    if (cfg->min_num_samples_per_block > cfg->max_num_samples_per_block) { return false; }
    if (cfg->max_num_lookahead_samples < cfg->max_num_samples_per_block) { return false; }
    return true;
}

/* per-length constants of a job: host libm pow() for the Welch divisor (lpc.c:259) */
struct LenConst { uint32_t n; double div, gain, scale; };
struct LenCache {
    LenConst slot[8]; int used = 0;
    const LenConst &get(uint32_t n)
    {
        for (int i = 0; i < used; i++) { if (slot[i].n == n) { return slot[i]; } }
        LenConst lc; lc.n = n;
        lc.div = 4.0 * std::pow((double)(n - 1), -2.0);
        { const double m = (double)n - 1; lc.gain = (15 * (m - 1) * (m - 1) * (m - 1)) / (8 * m * (m - 2) * (m * m - 2 * m + 2)); }   /* lpc.c:275-290 */
        lc.scale = 2.0 / n;
        const int at = (used < 8) ? used++ : (int)(n & 7u);
        slot[at] = lc;
        return slot[at];
    }
};

inline Job make_job(uint32_t stream, uint32_t offset, uint32_t n, uint32_t flags, LenCache &cache)
{
    Job j; const LenConst &lc = cache.get(n);
    j.stream = stream; j.offset = offset; j.nsmpl = n; j.flags = flags;
    j.welch_div = lc.div; j.welch_gain = lc.gain; j.ac_scale = lc.scale;
    return j;
}

uint32_t ceil_pow2_host(uint32_t v) { uint32_t p = 1; while (p < v) { p <<= 1; } return p; }

/* host-side view of one stream when PCM / output live in host memory */
struct HostStream { const void *ch[SRLA_MAX_NUM_CHANNELS]; };
struct HostIO {
    const HostStream *streams = nullptr;   /* per-channel host pointers of every stream */
    uint8_t *out = nullptr;                /* host output buffer */
    uint64_t out_capacity = 0;
    bool narrow = false;                   /* host samples are int32_t but the device layout is int16_t: they are
                                              narrowed by the feeder threads into pinned staging on their way in */
    const void *const *raw = nullptr;      /* WAV ingest: per stream, the interleaved frames of its data chunk (then
                                              `streams` is unused); they are copied as they are and de-interleaved on
                                              the device (Plan::raw_dev says where) */
};

/* Host feeder: the reference API hands over planar int32_t PCM in pageable memory.  For sources of
 * at most 16 bits a team of host threads narrows it to int16_t into a pinned staging buffer, group by group, while
 * the copy engine and the kernels work on the groups already staged: half the PCIe bytes, true asynchronous
 * copies, 2-byte kernel loads.  A sample outside the int16 range (a caller breaking the bits_per_sample
 * contract) is detected and the call falls back to the int32 layout. */
/* int32 -> int16 with a range check; nonzero when a sample does not fit.  The AVX2 version writes with streaming
 * stores (the staging buffer is only ever read by the copy engine) and packs with saturation: a chunk that saturates
 * is reported and the whole call falls back to the int32 layout, so its staging bytes are never used. */
static uint32_t narrow_scalar(const int32_t *src, int16_t *dst, uint32_t n)
{
    uint32_t bad = 0;
    for (uint32_t k = 0; k < n; k++) { const int32_t v = src[k]; bad |= (uint32_t)(v + 32768) >> 16; dst[k] = (int16_t)v; }
    return bad;
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static uint32_t narrow_avx2(const int32_t *src, int16_t *dst, uint32_t n)
{
    uint32_t i = 0, bad = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31u)) { const int32_t v = src[i]; bad |= (uint32_t)(v + 32768) >> 16; dst[i] = (int16_t)v; i++; }
    __m256i acc = _mm256_setzero_si256();
    const __m256i bias = _mm256_set1_epi32(32768);
    for (; i + 16u <= n; i += 16u) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 8));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_srli_epi32(_mm256_add_epi32(a, bias), 16), _mm256_srli_epi32(_mm256_add_epi32(b, bias), 16)));
        const __m256i packed = _mm256_permute4x64_epi64(_mm256_packs_epi32(a, b), 0xD8);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), packed);
    }
    _mm_sfence();
    if (!_mm256_testz_si256(acc, acc)) { bad |= 1u; }
    for (; i < n; i++) { const int32_t v = src[i]; bad |= (uint32_t)(v + 32768) >> 16; dst[i] = (int16_t)v; }
    return bad;
}
static uint32_t narrow_chunk(const int32_t *src, int16_t *dst, uint32_t n)
{
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    return have_avx2 ? narrow_avx2(src, dst, n) : narrow_scalar(src, dst, n);
}
#else
static uint32_t narrow_chunk(const int32_t *src, int16_t *dst, uint32_t n) { return narrow_scalar(src, dst, n); }
#endif

struct Feeder {
    struct Chunk { const int32_t *src; int16_t *dst; uint32_t count; uint32_t group; };
    std::vector<Chunk> chunks;
    std::unique_ptr<std::atomic<int>[]> left;      /* chunks of each group still to convert */
    std::atomic<size_t> next{0};
    std::atomic<int> overflow{0};
    /* output phase: once the input is staged the same threads copy the encoded bytes from the pinned staging
     * buffer (the target of the asynchronous D2H copies) into the caller's pageable buffer, 1 MB at a time, as soon
     * as the bytes are there */
    static constexpr unsigned long long kOutChunk = 1ull << 20, kUnknown = ~0ull;
    const uint8_t *out_stage = nullptr; uint8_t *out_user = nullptr;
    std::atomic<unsigned long long> out_ready{0}, out_total{kUnknown}, out_next{0};
    std::atomic<int> abort{0};
    WorkerPool *pool = nullptr;
    void start(size_t num_groups, int threads, WorkerPool *workers)
    {
        left.reset(new std::atomic<int>[num_groups]);
        for (size_t g = 0; g < num_groups; g++) { left[g].store(0); }
        for (const Chunk &ck : chunks) { left[ck.group].fetch_add(1); }
        pool = workers;
        pool->ensure(threads);
        pool->start([this] { work(); });
    }
    void work()
    {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= chunks.size()) { break; }
            const Chunk &ck = chunks[i];
            if (narrow_chunk(ck.src, ck.dst, ck.count)) { overflow.store(1); }
            left[ck.group].fetch_sub(1, std::memory_order_release);
        }
        if (out_stage == nullptr) { return; }
        for (;;) {
            const unsigned long long begin = out_next.fetch_add(1) * kOutChunk;
            for (;;) {
                if (abort.load(std::memory_order_acquire)) { return; }
                const unsigned long long total = out_total.load(std::memory_order_acquire);
                if (total != kUnknown && begin >= total) { return; }
                const unsigned long long want = std::min(begin + kOutChunk, total);          /* total == kUnknown: the full chunk */
                if (out_ready.load(std::memory_order_acquire) >= want) { std::memcpy(out_user + begin, out_stage + begin, (size_t)(want - begin)); break; }
                std::this_thread::yield();
            }
        }
    }
    template <typename Poll>
    void wait_group(size_t g, Poll poll) { while (left[g].load(std::memory_order_acquire) > 0) { poll(); std::this_thread::yield(); } }
    void join() { if (pool) { pool->wait(); pool = nullptr; } }
    ~Feeder() { abort.store(1, std::memory_order_release); join(); }
};

/* what one call encodes */
struct Plan {
    const struct SRLAB200Stream *streams = nullptr;   /* DEVICE PCM layout of every stream */
    uint32_t num_streams = 0;
    uint32_t nch = 0;
    bool emit_stream_header = true;
    bool use_fixed_lshift = false;
    uint32_t fixed_lshift = 0;
    bool variable = false;            /* min != max: optimal block division */
    bool size_only = false;           /* ComputeBlockSize */
    bool want_diag = false;
    bool allow_pipeline = true;
    const void *const *raw_dev = nullptr;   /* WAV ingest: per stream, device address of its interleaved frames */
    uint32_t container_bytes = 0;           /* bytes per sample there */
};

struct Runner {
    SRLAEncoder *enc; DeviceCtx *c;
    LenCache len_cache;
    uint64_t launches = 0;
    bool narrow_overflow = false;      /* the feeder met a sample outside int16: the caller redoes the call with the int32 layout */
    /* pipelined host path: called on the caller's thread whenever more encoded bytes have arrived in host memory --
     * (where they can be read, how many of the stream's bytes are there).  EncodeWhole reports progress through it while
     * the later groups are still being encoded (srla_encoder.c:1780-1782 calls back after every block). */
    std::function<void(const uint8_t *, uint64_t)> on_ready;
    bool serial_streams = false;       /* odd block size: the analysis calls of a stream form one chain (front_big_kernel) */

    LaunchParams base_params(const Plan &pl, uint32_t nmax) const
    {
        LaunchParams p;
        std::memset(&p, 0, sizeof(p));
        p.streams = (StreamDev *)c->streams.p;
        p.num_streams = pl.num_streams;
        p.nch = pl.nch; p.ncand = (pl.nch >= 2) ? pl.nch + 2 : pl.nch;
        p.bps = enc->param.bits_per_sample;
        p.max_order = enc->max_order;
        p.ltp_order = enc->param.ltp_order;
        p.nmax = nmax; p.fft_max = ceil_pow2_host(nmax);
        p.res_stride = round_up_u32(nmax, 4);
        p.sampling_rate = enc->param.sampling_rate; p.max_block = enc->param.max_num_samples_per_block; p.preset = enc->param.preset;
        p.fixed_lshift = pl.fixed_lshift; p.use_fixed_lshift = pl.use_fixed_lshift ? 1u : 0u;
        p.emit_stream_header = pl.emit_stream_header ? 1u : 0u;
        p.unit = std::ldexp(1.0, -(int)(p.bps - 1));
        p.tw_complex = (const double2 *)c->tw_complex.p; p.tw_real = (const double2 *)c->tw_real.p;
        for (int i = 0; i < 20; i++) { p.tw_complex_off[i] = c->tw_c_off[i]; p.tw_real_off[i] = c->tw_r_off[i]; }
        p.rice_threshold = (const double *)c->rice_thr.p;
        p.huff_code = (const uint32_t *)c->huff_code.p; p.huff_len = (const uint8_t *)c->huff_len.p;
        p.running = (unsigned long long *)c->misc.p;
        p.stats = (uint32_t *)((unsigned char *)c->misc.p + 2 * sizeof(unsigned long long));
        p.stream_begin = (unsigned long long *)((unsigned char *)c->misc.p + kMiscBytes);       /* behind the counters: one result copy */
        return p;
    }

    DeviceCtx::Lane &lane_of(cudaStream_t on)
    {
        for (int l = 1; l < kMaxLanes; l++) { if (c->lane[l].stream == on) { return c->lane[l]; } }
        return c->lane[0];
    }

    bool mark(size_t batch, int slot, cudaStream_t on)
    {
        const size_t idx = batch * 5 + (size_t)slot;
        while (c->ev_pool.size() <= idx) { cudaEvent_t e; CU_TRY(cudaEventCreate(&e)); c->ev_pool.push_back(e); }
        CU_TRY(cudaEventRecord(c->ev_pool[idx], on));
        return true;
    }

    /* dynamic shared memory opt-in + carve-out, once per (kernel, size).  `ctas` = CTAs per SM the kernel is
     * meant to run with: the carve-out is sized for exactly that many (plus the 1 KB the driver reserves per
     * CTA) so that the rest of the 228 KB stays L1 -- the FFT twiddle tables live there.  0 = maximum carve-out. */
    template <typename K>
    bool prep_kernel(K kernel, uint32_t smem_bytes, int ctas = 0)
    {
        /* the attributes are state of the FUNCTION on a device, shared by every handle of the process: the cache is
         * process-wide, and the dynamic shared-memory limit is only ever raised */
        std::lock_guard<std::mutex> lock(g_func_attr.m);
        FuncAttr &fa = g_func_attr.set[std::make_pair(c->device, reinterpret_cast<const void *>(kernel))];
        if (smem_bytes > fa.max_dynamic) {
            CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            fa.max_dynamic = smem_bytes;
        }
        int carve = cudaSharedmemCarveoutMaxShared;
        if (ctas > 0 && c->sized_carveout) {
            const double want = (double)ctas * (smem_bytes + 1024.0 + 256.0) / (228.0 * 1024.0) * 100.0;
            carve = (int)std::min(100.0, std::ceil(want));
        }
        if (carve != fa.carveout) {
            CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            fa.carveout = carve;
        }
        return true;
    }

    uint32_t svr_grid(const LaunchParams &p) const
    {
        const SvrLayout SL = make_svr_layout(p.nmax, p.max_order);
        const uint32_t per_sm = std::max(1u, std::min(8u, (uint32_t)(227u * 1024u) / (SL.total + 1024u)));
        return std::min(p.num_jobs * p.ncand, (uint32_t)c->num_sms * per_sm);
    }

    /* the three analysis kernels: front (autocorrelation) -> lpc (Levinson-Durbin) -> residual (FIR + Rice search) */
    /* tails [tail_lo, tail_hi) of c->tails_scratch lie in this launch's jobs, which start at index group_first of c->jobs */
    bool launch_analyse(const LaunchParams &p_in, size_t batch, cudaStream_t on, size_t tail_lo = 0, size_t tail_hi = 0, uint32_t group_first = 0,
                        bool pcm16_only = false, const Job *tail_list = nullptr)
    {
        LaunchParams p = p_in;
        const FrontLayout FL = make_front_layout(p.nmax, p.fft_max, p.ltp_order);
        const LpcLayout LL = make_lpc_layout(p.max_order);
        const ResidLayout RL = make_resid_layout(p.nmax, p.max_order);
        const uint32_t ncands = p.num_jobs * p.ncand;
        const dim3 grid(ncands), block(kThreads);
        const bool ltp = p.ltp_order > 0u;
        /* blocks whose transform / signal / residual do not fit an SM's shared memory (more than kMaxSharedBlock samples):
         * the same stages on per-CTA scratch in global memory, persistent CTAs */
        const bool big = p.serial_streams || p.nmax > (uint32_t)kMaxSharedBlock || (int)std::max(FL.total, RL.total) + 1024 > c->max_smem_optin;
        if (big) {
            if (p.svr_iterations > 0u) { std::fprintf(stderr, "[srla_b200] SVR refinement is limited to blocks of at most %d samples\n", kMaxSharedBlock); return false; }
            const FrontBigLayout FB = make_front_big_layout(p.nmax, p.fft_max);
            const uint32_t front_ctas = std::min(p.serial_streams ? std::max(1u, p.num_streams) : p.num_jobs, (uint32_t)c->num_sms * 2u);
            const uint32_t resid_ctas = std::min(ncands, (uint32_t)c->num_sms * 4u);
            const size_t need = std::max((size_t)front_ctas * FB.total, (size_t)resid_ctas * round_up_u32(RL.total, 256));
            if (!lane_of(on).big.reserve(need)) { return false; }
            LaunchParams pb = p;
            pb.big_scratch = (unsigned char *)lane_of(on).big.p; pb.big_stride = FB.total;
            if (ltp) { front_big_kernel<true><<<front_ctas, 1024, 0, on>>>(pb); } else { front_big_kernel<false><<<front_ctas, 1024, 0, on>>>(pb); }
            launches++;
            if (!mark(batch, 1, on)) { return false; }
            if (p.max_order > 0) {
                if (!prep_kernel(lpc_levinson_kernel, LL.total) || !prep_kernel(lpc_select_kernel, LL.select_total)) { return false; }
                lpc_levinson_kernel<<<(ncands + 31u) / 32u, 32, LL.total, on>>>(p);
                lpc_select_kernel<<<(ncands + 31u) / 32u, 128, LL.select_total, on>>>(p);
                launches += 2;
            }
            if (!mark(batch, 2, on)) { return false; }
            pb.big_stride = round_up_u32(RL.total, 256);
            residual_big_kernel<<<resid_ctas, block, 0, on>>>(pb);
            launches++;
            CU_TRY(cudaGetLastError());
            return true;
        }
        const Front16Layout F16 = make_front16_layout(p.nmax, p.fft_max, p.nch);
        if (p.fft_max <= 4096u && !ltp && pcm16_only && p.nch <= 2u && c->front16 && p.max_order > 0u) {
            /* 16-bit PCM without LTP: one CTA per job, the rows of all channels staged once by bulk asynchronous copies */
            p.f16 = F16;
            if (!prep_kernel(front16_kernel<128>, F16.total, 3)) { return false; }
            front16_kernel<128><<<p.num_jobs, 128, F16.total, on>>>(p);
        } else if (p.fft_max <= 4096u) {
            /* 128 threads: every thread owns one 16-point FFT work unit (2048 complex points / 16) */
            if (ltp) {
                if (!prep_kernel(front_kernel<128, 3, true>, FL.total, 3)) { return false; }
                front_kernel<128, 3, true><<<grid, 128, FL.total, on>>>(p);
            } else if (c->front_occ == 2) {
                if (!prep_kernel(front_kernel<128, 2, false>, FL.total, 2)) { return false; }
                front_kernel<128, 2, false><<<grid, 128, FL.total, on>>>(p);
            } else if (c->front_occ == 4) {
                if (!prep_kernel(front_kernel<128, 4, false>, FL.total, 4)) { return false; }
                front_kernel<128, 4, false><<<grid, 128, FL.total, on>>>(p);
            } else {
                if (!prep_kernel(front_kernel<128, 3, false>, FL.total, 3)) { return false; }
                front_kernel<128, 3, false><<<grid, 128, FL.total, on>>>(p);
            }
        } else if (p.fft_max <= 8192u) {
            if (ltp) {
                if (!prep_kernel(front_kernel<256, 2, true>, FL.total, 2)) { return false; }
                front_kernel<256, 2, true><<<grid, 256, FL.total, on>>>(p);
            } else {
                if (!prep_kernel(front_kernel<256, 2, false>, FL.total, 2)) { return false; }
                front_kernel<256, 2, false><<<grid, 256, FL.total, on>>>(p);
            }
        } else if (p.fft_max <= 16384u) {
            /* one CTA of 512 threads per SM: the transform (128 KB) and the signal (64 KB) fill its shared memory */
            if (ltp) {
                if (!prep_kernel(front_kernel<512, 1, true>, FL.total, 1)) { return false; }
                front_kernel<512, 1, true><<<grid, 512, FL.total, on>>>(p);
            } else {
                if (!prep_kernel(front_kernel<512, 1, false>, FL.total, 1)) { return false; }
                front_kernel<512, 1, false><<<grid, 512, FL.total, on>>>(p);
            }
        } else {
            std::fprintf(stderr, "[srla_b200] block of %u samples exceeds the pipeline capacity (%d)\n", p.nmax, kMaxBlock);
            return false;
        }
        launches++;
        if (tail_hi > tail_lo) {
            /* the reference's stale-scratch corners on the last block of a stream (see front_tail_kernel) */
            const uint32_t pbuf_len = std::max(p.fft_max, 512u);
            const uint32_t pbuf_bytes = 8u * (pbuf_len + 272u);
            const TailJob *d_tails = (const TailJob *)c->tails.p + tail_lo;
            const uint32_t nt = (uint32_t)(tail_hi - tail_lo);
            /* the replica of the reference's scratch buffer sits behind the kernel's shared memory when it fits, else in global memory */
            const bool in_smem = (int)(FL.total + pbuf_bytes) + 1024 <= c->max_smem_optin;
            const uint32_t smem_tail = in_smem ? FL.total + pbuf_bytes : FL.total;
            double *pbuf_global = nullptr;
            if (!in_smem) {
                if (!c->tail_scratch.reserve((size_t)nt * pbuf_bytes)) { return false; }
                pbuf_global = (double *)c->tail_scratch.p;
            }
            const Job *all = tail_list ? tail_list : (const Job *)c->jobs.p;       /* the call list the tails' predecessors are looked up in */
            if (p.fft_max <= 4096u) {
                if (ltp) { if (!prep_kernel(front_tail_kernel<128, true>, smem_tail)) { return false; } front_tail_kernel<128, true><<<nt, 128, smem_tail, on>>>(p, d_tails, all, group_first, pbuf_len, pbuf_global); }
                else { if (!prep_kernel(front_tail_kernel<128, false>, smem_tail)) { return false; } front_tail_kernel<128, false><<<nt, 128, smem_tail, on>>>(p, d_tails, all, group_first, pbuf_len, pbuf_global); }
            } else if (p.fft_max <= 8192u) {
                if (ltp) { if (!prep_kernel(front_tail_kernel<256, true>, smem_tail)) { return false; } front_tail_kernel<256, true><<<nt, 256, smem_tail, on>>>(p, d_tails, all, group_first, pbuf_len, pbuf_global); }
                else { if (!prep_kernel(front_tail_kernel<256, false>, smem_tail)) { return false; } front_tail_kernel<256, false><<<nt, 256, smem_tail, on>>>(p, d_tails, all, group_first, pbuf_len, pbuf_global); }
            } else {
                if (ltp) { if (!prep_kernel(front_tail_kernel<512, true>, smem_tail)) { return false; } front_tail_kernel<512, true><<<nt, 512, smem_tail, on>>>(p, d_tails, all, group_first, pbuf_len, pbuf_global); }
                else { if (!prep_kernel(front_tail_kernel<512, false>, smem_tail)) { return false; } front_tail_kernel<512, false><<<nt, 512, smem_tail, on>>>(p, d_tails, all, group_first, pbuf_len, pbuf_global); }
            }
            launches++;
        }
        if (!mark(batch, 1, on)) { return false; }
        if (p.max_order > 0) {
            if (!prep_kernel(lpc_levinson_kernel, LL.total) || !prep_kernel(lpc_select_kernel, LL.select_total)) { return false; }
            lpc_levinson_kernel<<<(ncands + 31u) / 32u, 32, LL.total, on>>>(p);
            lpc_select_kernel<<<(ncands + 31u) / 32u, 128, LL.select_total, on>>>(p);
            launches += 2;
            if (p.svr_iterations > 0u) {
                /* persistent CTAs: one covariance / Cholesky matrix of P x P doubles each in global memory */
                const SvrLayout SL = make_svr_layout(p.nmax, p.max_order);
                if (!prep_kernel(svr_kernel, SL.total)) { return false; }
                svr_kernel<<<svr_grid(p), kThreads, SL.total, on>>>(p);
                launches++;
            }
        }
        if (!mark(batch, 2, on)) { return false; }
        if (p.ltp_order == 0u && pcm16_only && c->resid16) {
            /* 16-bit PCM without LTP: persistent CTAs, source rows staged one item ahead by bulk asynchronous copies */
            const Resid16Layout R16 = make_resid16_layout(p.nmax, p.max_order);
            if (!prep_kernel(residual16_kernel, R16.total)) { return false; }
            const uint32_t per_sm = std::max(1u, std::min((uint32_t)SRLA_R16_OCC, (uint32_t)(227u * 1024u) / (R16.total + 1024u)));
            residual16_kernel<<<std::min(ncands, (uint32_t)c->num_sms * per_sm), block, R16.total, on>>>(p);
        } else {
            if (!prep_kernel(residual_kernel, RL.total)) { return false; }
            residual_kernel<<<grid, block, RL.total, on>>>(p);
        }
        launches++;
        CU_TRY(cudaGetLastError());
        return true;
    }

    /* upload stream descriptors (or_mask = 0, lshift = 0) */
    bool prepare_streams(const Plan &pl)
    {
        const size_t bytes = sizeof(StreamDev) * pl.num_streams;
        if (!c->streams.reserve(bytes) || !c->h_small.reserve(bytes + 64)) { return false; }
        StreamDev *h = (StreamDev *)c->h_small.p;
        for (uint32_t s = 0; s < pl.num_streams; s++) {
            h[s].pcm = pl.streams[s].pcm; h[s].stride = pl.streams[s].channel_stride;
            h[s].num_samples = pl.streams[s].num_samples; h[s].sample_bytes = pl.streams[s].sample_bytes;
            h[s].lshift = 0; h[s].or_mask = 0;
            h[s].raw = pl.raw_dev ? pl.raw_dev[s] : nullptr; h[s].container_bytes = pl.container_bytes; h[s].pad_ = 0;
        }
        CU_TRY(cudaMemcpyAsync(c->streams.p, h, bytes, cudaMemcpyHostToDevice, c->stream));
        return true;
    }

    /* OR-reduce the samples of `count` jobs into their streams and refresh every stream's shift */
    /* with `ingest` the jobs' frames have just arrived as interleaved WAV data: they are de-interleaved first, and
     * that kernel ORs the samples on its way */
    bool launch_lshift(const Plan &pl, const Job *d_jobs, uint32_t count, uint32_t *d_snapshot, cudaStream_t on, bool ingest = false)
    {
        if (ingest) { deinterleave_jobs_kernel<<<count, 256, 0, on>>>((StreamDev *)c->streams.p, d_jobs, pl.nch); }
        else { lshift_jobs_kernel<<<std::min(count, (uint32_t)c->num_sms * 16u), 128, 0, on>>>((StreamDev *)c->streams.p, d_jobs, pl.nch, count); }
        lshift_finish_kernel<<<(pl.num_streams + 255) / 256, 256, 0, on>>>((StreamDev *)c->streams.p, pl.num_streams, d_snapshot);
        CU_TRY(cudaGetLastError());
        launches += 2;
        return true;
    }

    /* analyse + decide a list of jobs (already on the device at d_jobs), optionally scan + emit, on lane `ln`.
     * scan_after: event the output-offset scan has to wait for (the previous group's scan, on another lane);
     * scan_done: recorded once this group's scan has run. */
    bool run_batch(const Plan &pl, const Job *d_jobs, uint32_t count, uint32_t nmax, bool emit, uint8_t *d_out, uint64_t cap,
                   bool store_residual, size_t ev_idx, unsigned long long *h_mailbox, int ln = 0,
                   cudaEvent_t scan_after = nullptr, cudaEvent_t scan_done = nullptr, bool with_tails = false, const Job *tail_list = nullptr)
    {
        DeviceCtx::Lane &L = c->lane[ln];
        const cudaStream_t on = L.stream;
        LaunchParams p = base_params(pl, nmax);
        p.jobs = d_jobs; p.num_jobs = count;
        const size_t ncand = p.ncand;
        if (!L.cand.reserve(sizeof(CandOut) * ncand * count) || !L.jobout.reserve(sizeof(JobOut) * count)) { return false; }
        if (store_residual && !L.residual.reserve(sizeof(int32_t) * ncand * count * (size_t)p.res_stride)) { return false; }
        if (pl.want_diag && !L.diag.reserve(sizeof(CandDiag) * ncand * count)) { return false; }
        p.lag_stride = round_up_u32(p.max_order + 2u, 2);
        if (!L.lags.reserve(sizeof(double) * round_up_u32((uint32_t)(ncand * count), 32) * (size_t)p.lag_stride)) { return false; }
        if (!L.lpc_state.reserve(sizeof(double) * (size_t)(round_up_u32((uint32_t)(ncand * count), 32) / 32u) * 2u * (p.max_order + 2u) * 32u)) { return false; }
        p.lags = (double *)L.lags.p; p.lpc_state = (double *)L.lpc_state.p;
        p.svr_iterations = enc->param.num_svr_filter_learning_iteration;
        if (p.svr_iterations > 0u && p.max_order > 0u) {
            if (!L.svr_coef.reserve(sizeof(double) * (size_t)ncand * count * p.max_order)
                || !L.svr_matrix.reserve(sizeof(double) * (size_t)svr_grid(p) * p.max_order * p.max_order)) { return false; }
            p.svr_coef = (double *)L.svr_coef.p; p.svr_matrix = (double *)L.svr_matrix.p;
        }
        p.cand = (CandOut *)L.cand.p; p.jobout = (JobOut *)L.jobout.p;
        p.residual = store_residual ? (int32_t *)L.residual.p : nullptr;
        p.diag = pl.want_diag ? (CandDiag *)L.diag.p : nullptr;
        p.out = d_out; p.out_capacity = cap;
        const uint32_t raw_max = 11u + (uint32_t)(((uint64_t)p.bps * nmax * p.nch) / 8u);
        p.emit_smem_bytes = raw_max;
        if (!mark(ev_idx, 0, on)) { return false; }
        size_t tail_lo = 0, tail_hi = 0; uint32_t group_first = 0;
        if (with_tails) {
            /* d_jobs points into the call's fixed tiling (c->jobs): the tails whose job lies in [group_first, group_first + count) */
            group_first = (uint32_t)(d_jobs - (const Job *)c->jobs.p);
            const std::vector<TailJob> &tv = c->tails_scratch;
            tail_lo = std::lower_bound(tv.begin(), tv.end(), group_first, [](const TailJob &t, uint32_t v) { return t.out < v; }) - tv.begin();
            tail_hi = std::lower_bound(tv.begin(), tv.end(), group_first + count, [](const TailJob &t, uint32_t v) { return t.out < v; }) - tv.begin();
        }
        bool pcm16_only = true;
        for (uint32_t s = 0; s < pl.num_streams && pcm16_only; s++) { pcm16_only = pl.streams[s].sample_bytes == 2u; }
        p.replay_tails = (with_tails && !pl.variable) ? 1u : 0u; p.group_first = group_first; p.jobs_all = (const Job *)c->jobs.p;     /* big-block path */
        p.serial_streams = (serial_streams && !pl.variable) ? 1u : 0u;
        if (!launch_analyse(p, ev_idx, on, tail_lo, tail_hi, group_first, pcm16_only, tail_list)) { return false; }
        if (!mark(ev_idx, 3, on)) { return false; }
        decide_kernel<<<(count + 127) / 128, 128, 0, on>>>(p);
        launches++;
        if (emit) {
            if (scan_after) { CU_TRY(cudaStreamWaitEvent(on, scan_after, 0)); }
            scan_kernel<<<1, 1024, 0, on>>>(p);
            if (h_mailbox) { CU_TRY(cudaMemcpyAsync(h_mailbox, p.running, sizeof(unsigned long long), cudaMemcpyDeviceToHost, on)); }
            if (scan_done) { CU_TRY(cudaEventRecord(scan_done, on)); }      /* after the mailbox copy: the next scan overwrites running[0] */
            const uint32_t smem = round_up_u32(raw_max, 4) + 16u;
            if ((int)smem + 2048 > c->max_smem_optin) {
                /* a block larger than an SM's shared memory is staged in global memory by persistent CTAs */
                const uint32_t ctas = std::min(count, (uint32_t)c->num_sms * 4u);
                LaunchParams pb = p;
                pb.big_stride = round_up_u32(smem, 256);
                if (!L.big.reserve((size_t)ctas * pb.big_stride)) { return false; }
                pb.big_scratch = (unsigned char *)L.big.p;
                emit_big_kernel<<<ctas, kThreads, 0, on>>>(pb);
            } else {
                if (!prep_kernel(emit_kernel, smem)) { return false; }
                emit_kernel<<<count, kThreads, smem, on>>>(p);
            }
            launches += 2;
        }
        CU_TRY(cudaGetLastError());
        if (!mark(ev_idx, 4, on)) { return false; }
        return true;
    }

    uint32_t jobs_per_batch(const Plan &pl, uint32_t nmax) const
    {
        const size_t ncand = (pl.nch >= 2) ? pl.nch + 2 : pl.nch;
        const size_t per_job = sizeof(int32_t) * ncand * round_up_u32(nmax, 4) + sizeof(CandOut) * ncand;
        size_t budget = (size_t)768 << 20;
        size_t n = budget / per_job;
        return (uint32_t)std::max<size_t>(64, std::min<size_t>(n, 65536));
    }

    /* the reference's block division of one look-ahead chunk from exact candidate sizes
     * (srla_encoder.c:249-307 dense shortest path, 310-424 graph construction / back trace) */
    static void shortest_partition(const std::vector<uint32_t> &edge_bytes /* [nodes*nodes], 0 = no edge */, uint32_t nodes,
                                   uint32_t unit, uint32_t n, std::vector<uint32_t> &parts)
    {
        const double BIG = (double)(1UL << 24);
        std::vector<double> dist(nodes, BIG); std::vector<uint32_t> from(nodes, ~0u); std::vector<uint8_t> done(nodes, 0);
        dist[0] = 0.0;
        uint32_t cur = 0;
        for (;;) {
            double low = BIG;
            for (uint32_t i = 0; i < nodes; i++) { if (!done[i] && low > dist[i]) { low = dist[i]; cur = i; } }
            if (cur == nodes - 1) { break; }
            for (uint32_t i = 0; i < nodes; i++) {
                const uint32_t eb = edge_bytes[cur * nodes + i];
                const double w = eb ? (double)eb : BIG;
                if (dist[i] > w + dist[cur]) { dist[i] = w + dist[cur]; from[i] = cur; }
            }
            done[cur] = 1;
        }
        std::vector<uint32_t> rev;
        for (uint32_t node = nodes - 1; node != 0; node = from[node]) {
            uint32_t len = (node - from[node]) * unit;
            if (len > n - from[node] * unit) { len = n - from[node] * unit; }
            rev.push_back(len);
        }
        parts.assign(rev.rbegin(), rev.rend());
    }

    bool upload_jobs(const std::vector<Job> &jobs)
    {
        const size_t bytes = sizeof(Job) * jobs.size();
        if (c->upload_pending) { CU_TRY(cudaEventSynchronize(c->ev_upload)); c->upload_pending = false; }   /* the staging buffer is reused */
        c->jobs_cached = false;                              /* the device copy is being replaced */
        if (!c->jobs.reserve(bytes) || !c->h_jobs.reserve(bytes)) { return false; }
        std::memcpy(c->h_jobs.p, jobs.data(), bytes);
        CU_TRY(cudaMemcpyAsync(c->jobs.p, c->h_jobs.p, bytes, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaEventRecord(c->ev_upload, c->stream));
        c->upload_pending = true;
        return true;
    }

    /* host -> device copy of the samples [begin, end) of every channel of stream s */
    bool h2d_range(const Plan &pl, const HostIO &io, uint32_t s, uint32_t begin, uint32_t end, cudaStream_t on)
    {
        const struct SRLAB200Stream &d = pl.streams[s];
        const size_t sb = d.sample_bytes;
        if (io.raw) {
            const size_t frame = (size_t)pl.nch * pl.container_bytes;
            CU_TRY(cudaMemcpyAsync((unsigned char *)pl.raw_dev[s] + begin * frame, (const unsigned char *)io.raw[s] + begin * frame,
                                   (size_t)(end - begin) * frame, cudaMemcpyHostToDevice, on));
            return true;
        }
        for (uint32_t ch = 0; ch < pl.nch; ch++) {
            unsigned char *dst = (unsigned char *)d.pcm + ((size_t)d.channel_stride * ch + begin) * sb;
            /* narrowed input: the staging buffer mirrors the device layout byte for byte */
            const unsigned char *src = io.narrow ? (const unsigned char *)c->h_stage.p + (dst - (unsigned char *)c->pcm.p)
                                                 : (const unsigned char *)io.streams[s].ch[ch] + (size_t)begin * sb;
            CU_TRY(cudaMemcpyAsync(dst, src, (size_t)(end - begin) * sb, cudaMemcpyHostToDevice, on));
        }
        return true;
    }

    /* whole call: streams -> output.
     * io == NULL: PCM is resident at pl.streams[].pcm and the output stays in d_out.
     * io != NULL: PCM is copied from host memory into the layout pl.streams[] describes and the output
     *   is copied back to io->out.  With fixed blocks the work is split into groups whose H2D copy,
     *   kernels and D2H copy overlap on three streams.  offset_lshift needs the whole stream, so a
     *   group runs with the shift of the samples seen SO FAR (snapshotted per group); if a later
     *   group lowers a stream's shift -- essentially never for real audio, whose first odd sample is
     *   in the first block -- the call is redone unpipelined on the now resident data. */
    SRLAApiResult run(const Plan &pl, uint8_t *d_out, uint64_t cap, uint64_t *stream_offsets, uint32_t *single_estimate, const HostIO *io)
    {
        SRLAB200Stats &stt = enc->stats;
        std::memset(&stt, 0, sizeof(stt));
        if (cudaSetDevice(c->device) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        const uint32_t max_block = enc->param.max_num_samples_per_block, min_block = enc->param.min_num_samples_per_block;
        for (uint32_t s = 0; s < pl.num_streams; s++) {
            if (pl.streams[s].num_samples == 0) { return SRLA_APIRESULT_INVALID_FORMAT; }        /* srla_encoder.c:107 */
            if (pl.streams[s].sample_bytes != 2 && pl.streams[s].sample_bytes != 4) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
            if (pl.streams[s].pcm == nullptr) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
            stt.bytes_in += (uint64_t)pl.nch * pl.streams[s].num_samples * pl.streams[s].sample_bytes;
        }

        /* ---- job list ---- */
        /* repeated calls on equally shaped input (a service encoding batch after batch) reuse the tiling and its
         * device copy: building and uploading 10^4 jobs would leave the device idle between calls */
        std::vector<Job> &jobs = c->jobs_scratch;
        std::vector<uint32_t> key;
        key.reserve(pl.num_streams + 2);
        key.push_back(max_block); key.push_back(pl.num_streams); key.push_back(enc->param.ltp_order); key.push_back(enc->max_order);
        for (uint32_t s = 0; s < pl.num_streams; s++) { key.push_back(pl.streams[s].num_samples); }
        const bool reuse_jobs = c->jobs_cached && !pl.variable && key == c->jobs_key;
        if (!reuse_jobs) {
            c->jobs_cached = false;
            jobs.clear();
            c->tails_scratch.clear();
            /* stale-scratch corners of the reference (front_tail_kernel): only the last block of a stream can be odd or, with
             * LTP, shorter than the 263 lags the pitch search reads -- as long as the block size itself is neither */
            const uint32_t ltp = enc->param.ltp_order;
            const bool replay_tails = !pl.variable && (max_block & 1u) == 0u && (ltp == 0u || max_block >= 263u);
            for (uint32_t s = 0; s < pl.num_streams; s++) {          /* fixed tiling: the block list, or the cover used for the shift */
                const uint32_t total = pl.streams[s].num_samples;
                const uint32_t first = (uint32_t)jobs.size();
                for (uint32_t at = 0; at < total; at += max_block) {
                    jobs.push_back(make_job(s, at, std::min(max_block, total - at), at == 0 ? kJobFirstOfStream : 0u, len_cache));
                }
                const uint32_t last_n = jobs.back().nsmpl;
                if (replay_tails && last_n > enc->max_order && ((last_n & 1u) || (ltp > 0u && ceil_pow2_host(last_n) < 263u))) {
                    TailJob tj; tj.job = (uint32_t)jobs.size() - 1u; tj.first_of_stream = first; tj.out = tj.job;
                    c->tails_scratch.push_back(tj);
                }
            }
        }
        /* An ODD block size makes every block's Welch window keep a sample of the previous analysis call (lpc.c:260-264): the
         * calls of a stream form one chain, which front_big_kernel walks block by block -- all blocks of a stream in one launch */
        serial_streams = !pl.variable && (max_block & 1u) != 0u && (enc->param.ltp_order == 0u || max_block >= 263u) && enc->max_order > 0u;
        /* large fixed-block calls are split into groups that alternate between the lanes (compute streams);
         * with host I/O every group additionally has its own H2D / D2H copies on the copy streams */
        const bool split = !pl.variable && !pl.size_only && pl.allow_pipeline && !serial_streams && jobs.size() >= 2048 && (io || c->split_device > 1);
        const bool pipelined = split && io && !pl.use_fixed_lshift;
        const int lanes = split ? (io ? c->lanes : c->split_device) : 1;
        uint32_t per_batch = jobs_per_batch(pl, max_block) / (uint32_t)lanes;
        if (serial_streams) {
            if (jobs.size() > (size_t)per_batch * 8u) { std::fprintf(stderr, "[srla_b200] odd block size %u: %zu blocks exceed what one serial launch holds\n", max_block, jobs.size()); return SRLA_APIRESULT_NG; }
            per_batch = (uint32_t)std::max<size_t>(per_batch, jobs.size());
        }
        /* group boundaries (job indices).  With host I/O the first groups are small so the kernels start as soon
         * as a little PCM has arrived, and the last ones are small so little output is left to copy back when the
         * kernels finish; in between the groups are large enough to fill the machine. */
        std::vector<size_t> gstart;
        gstart.push_back(0);
        if (!split) {
            for (size_t at = per_batch; at < jobs.size(); at += per_batch) { gstart.push_back(at); }
        } else if (pipelined && c->ramp && jobs.size() >= 4096) {
            static const double kRampDefault[] = { 0.03, 0.06, 0.11, 0.15, 0.15, 0.15, 0.14, 0.11, 0.06, 0.04 };
            std::vector<double> ramp_spec(kRampDefault, kRampDefault + sizeof(kRampDefault) / sizeof(kRampDefault[0]));
            if (const char *e = std::getenv("SRLA_B200_RAMP_SPEC")) {             /* tuning: comma-separated group fractions */
                std::vector<double> v; const char *q = e;
                while (*q) { char *end = nullptr; const double f = std::strtod(q, &end); if (end == q) { break; } v.push_back(f); q = (*end == ',') ? end + 1 : end; }
                if (!v.empty()) { ramp_spec = v; }
            }
            double acc = 0.0;
            for (double f : ramp_spec) {
                acc += f;
                size_t at = std::min(jobs.size(), (size_t)(acc * (double)jobs.size() + 0.5));
                while (at - gstart.back() > per_batch) { gstart.push_back(gstart.back() + per_batch); }
                if (at > gstart.back() && at < jobs.size()) { gstart.push_back(at); }
            }
        } else {
            const size_t want_groups = io ? (size_t)c->groups : (size_t)lanes;              /* device-resident: one group per lane */
            const size_t group = std::min<size_t>(per_batch, std::max<size_t>(512u, (jobs.size() + want_groups - 1) / want_groups));
            for (size_t at = group; at < jobs.size(); at += group) { gstart.push_back(at); }
        }
        gstart.push_back(jobs.size());
        const size_t num_groups = gstart.size() - 1;
        c->lane[0].stream = c->stream;                     /* lane 0 is the caller's stream */
        if (io && io->narrow && !pipelined) { std::fprintf(stderr, "[srla_b200] internal: narrowed input outside the pipelined path\n"); return SRLA_APIRESULT_NG; }

        if (cudaEventRecord(c->ev_begin, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        const auto host_t0 = std::chrono::steady_clock::now();
        auto host_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
        std::vector<double> host_h2d, host_launch, host_d2h;
        if (!prepare_streams(pl)) { return SRLA_APIRESULT_NG; }
        if (!c->misc.reserve(kMiscBytes + sizeof(unsigned long long) * (pl.num_streams + 1) + 64)) { return SRLA_APIRESULT_NG; }
        if (cudaMemsetAsync(c->misc.p, 0, kMiscBytes, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        if (!reuse_jobs) {
            if (!upload_jobs(jobs)) { return SRLA_APIRESULT_NG; }
            if (!pl.variable && !c->tails_scratch.empty()) {
                const size_t bytes = sizeof(TailJob) * c->tails_scratch.size();
                if (!c->tails.reserve(bytes)) { return SRLA_APIRESULT_NG; }
                /* pageable source: the copy is staged by the driver before the call returns */
                if (cudaMemcpyAsync(c->tails.p, c->tails_scratch.data(), bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
            }
            if (!pl.variable) { c->jobs_key = key; c->jobs_cached = true; }
        }
        const bool use_tails = !pl.variable && !c->tails_scratch.empty();

        /* ---- host input ---- */
        std::vector<cudaEvent_t> &h2d_done = c->ev_h2d;
        Feeder feeder;
        /* copies group g's samples to the device (copy stream) and records h2d_done[g] */
        std::function<void()> poll_d2h;                     /* set once the groups run: copies finished groups back while the host waits */
        auto issue_h2d = [&](size_t g) -> bool {
            if (io->narrow) { feeder.wait_group(g, [&] { if (poll_d2h) { poll_d2h(); } }); }
            const size_t j0 = gstart[g], j1 = gstart[g + 1];
            size_t j = j0;
            while (j < j1) {                                   /* one contiguous sample range per stream touched */
                const uint32_t s = jobs[j].stream, begin = jobs[j].offset;
                size_t k = j;
                while (k + 1 < j1 && jobs[k + 1].stream == s) { k++; }
                if (!h2d_range(pl, *io, s, begin, jobs[k].offset + jobs[k].nsmpl, c->copy_stream)) { return false; }
                j = k + 1;
            }
            const bool ok = cudaEventRecord(h2d_done[g], c->copy_stream) == cudaSuccess;
            if (c->trace) { if (host_h2d.size() <= g) { host_h2d.resize(g + 1, 0.0); } host_h2d[g] = host_ms(); }
            return ok;
        };
        bool copies_async = true;      /* pinned source: the H2D of the next group is queued before this group's kernels */
        if (io) {
            if (pipelined) {
                while (h2d_done.size() < num_groups) { cudaEvent_t e; if (cudaEventCreateWithFlags(&e, c->trace ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) { return SRLA_APIRESULT_NG; } h2d_done.push_back(e); }
                if (io->narrow) {
                    for (size_t g = 0; g < num_groups; g++) {
                        size_t j = gstart[g];
                        while (j < gstart[g + 1]) {
                            const uint32_t s = jobs[j].stream, begin = jobs[j].offset;
                            size_t k = j;
                            while (k + 1 < gstart[g + 1] && jobs[k + 1].stream == s) { k++; }
                            const uint32_t end = jobs[k].offset + jobs[k].nsmpl;
                            const struct SRLAB200Stream &d = pl.streams[s];
                            for (uint32_t ch = 0; ch < pl.nch; ch++) {
                                int16_t *dst = (int16_t *)((unsigned char *)c->h_stage.p + ((unsigned char *)d.pcm - (unsigned char *)c->pcm.p)) + (size_t)d.channel_stride * ch;
                                const int32_t *src = (const int32_t *)io->streams[s].ch[ch];
                                for (uint32_t at = begin; at < end; at += 65536u) {
                                    Feeder::Chunk ck; ck.src = src + at; ck.dst = dst + at; ck.count = std::min(65536u, end - at); ck.group = (uint32_t)g;
                                    feeder.chunks.push_back(ck);
                                }
                            }
                            j = k + 1;
                        }
                    }
                    if (c->h_stage_out.reserve(std::min<uint64_t>(cap, io->out_capacity))) { feeder.out_stage = (const uint8_t *)c->h_stage_out.p; feeder.out_user = io->out; }
                    if (!c->pool) { c->pool.reset(new WorkerPool()); }
                    feeder.start(num_groups, c->feed_threads, c->pool.get());
                } else {
                    cudaPointerAttributes attr;
                    if (cudaPointerGetAttributes(&attr, io->raw ? io->raw[0] : io->streams[0].ch[0]) != cudaSuccess || attr.type != cudaMemoryTypeHost) { copies_async = false; (void)cudaGetLastError(); }
                }
                if (!issue_h2d(0)) { return SRLA_APIRESULT_NG; }
            } else {
                for (uint32_t s = 0; s < pl.num_streams; s++) { if (!h2d_range(pl, *io, s, 0, pl.streams[s].num_samples, c->stream)) { return SRLA_APIRESULT_NG; } }
            }
        }

        size_t ev_idx = 0;
        uint32_t nmax = 1;
        const bool ingest = io && io->raw;
        if (ingest && pl.use_fixed_lshift) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
        if (!pipelined && !pl.use_fixed_lshift) {
            /* exact offset_lshift of every stream before any analysis */
            if (!launch_lshift(pl, (const Job *)c->jobs.p, (uint32_t)jobs.size(), nullptr, c->stream, ingest)) { return SRLA_APIRESULT_NG; }
        }

        bool var_final_tails = false;
        if (pl.variable) {
            /* pass 1: exact size of every candidate segment of every look-ahead chunk */
            struct Chunk { uint32_t stream, at, len, nodes, first_job, num_jobs, final_begin, final_end; };
            std::vector<Chunk> chunks; std::vector<Job> cand_jobs;
            const uint32_t step = enc->param.num_lookahead_samples;
            /* The reference's stale-scratch corners (front_tail_kernel) with variable blocks: a candidate segment clipped at
             * the end of a stream can be odd (or, with LTP, shorter than the 263 lags the pitch search reads), and then its
             * analysis depends on the calls in front of it.  cand_jobs lists the segments of a chunk in the reference's
             * call order (srla_encoder.c:352-388), so the predecessors of such a segment are the list entries in front of
             * it; the blocks the chunk is finally coded with follow the same way behind the whole search (:1660-1690). */
            const uint32_t ltp_order = enc->param.ltp_order;
            const bool var_tails = (max_block & 1u) == 0u && (min_block & 1u) == 0u && (ltp_order == 0u || min_block >= 263u)
                                   && enc->max_order > 0u && max_block <= (uint32_t)kMaxSharedBlock;
            auto dependent = [&](uint32_t n) { return n > enc->max_order && ((n & 1u) || (ltp_order > 0u && ceil_pow2_host(n) < 263u)); };
            c->tails_scratch.clear();
            for (uint32_t s = 0; s < pl.num_streams; s++) {
                const uint32_t total = pl.streams[s].num_samples;
                for (uint32_t at = 0; at < total; at += step) {
                    Chunk ch; ch.stream = s; ch.at = at; ch.len = std::min(step, total - at);
                    ch.nodes = (ch.len + min_block - 1) / min_block + 1; ch.first_job = (uint32_t)cand_jobs.size();
                    for (uint32_t i = 0; i < ch.nodes; i++) {
                        for (uint32_t j = i + 1; j < ch.nodes; j++) {
                            uint32_t len = (j - i) * min_block;
                            if (len > max_block) { continue; }
                            if (len > ch.len - i * min_block) { len = ch.len - i * min_block; }
                            if (var_tails && dependent(len)) {
                                TailJob tj; tj.job = (uint32_t)cand_jobs.size(); tj.first_of_stream = ch.first_job; tj.out = tj.job;
                                c->tails_scratch.push_back(tj);
                            }
                            cand_jobs.push_back(make_job(s, at + i * min_block, len, 0u, len_cache));
                        }
                    }
                    ch.num_jobs = (uint32_t)cand_jobs.size() - ch.first_job; ch.final_begin = ch.final_end = 0;
                    chunks.push_back(ch);
                }
            }
            stt.num_analysed += cand_jobs.size();
            std::vector<uint32_t> est(cand_jobs.size());
            if (!upload_jobs(cand_jobs)) { return SRLA_APIRESULT_NG; }
            auto upload_tails = [&]() {
                if (c->tails_scratch.empty()) { return true; }
                const size_t bytes = sizeof(TailJob) * c->tails_scratch.size();
                if (!c->tails.reserve(bytes)) { return false; }
                /* pageable source: the copy is staged by the driver before the call returns */
                return cudaMemcpyAsync(c->tails.p, c->tails_scratch.data(), bytes, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
            };
            if (!upload_tails()) { return SRLA_APIRESULT_NG; }
            for (size_t b = 0; b < cand_jobs.size(); b += per_batch) {
                const uint32_t cnt = (uint32_t)std::min<size_t>(per_batch, cand_jobs.size() - b);
                if (!run_batch(pl, (const Job *)c->jobs.p + b, cnt, max_block, false, nullptr, 0, false, ev_idx++, nullptr, 0, nullptr, nullptr,
                               !c->tails_scratch.empty())) { return SRLA_APIRESULT_NG; }
                if (!c->h_jobout.reserve(sizeof(JobOut) * cnt)) { return SRLA_APIRESULT_NG; }
                if (cudaMemcpyAsync(c->h_jobout.p, c->lane[0].jobout.p, sizeof(JobOut) * cnt, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess
                    || cudaStreamSynchronize(c->stream) != cudaSuccess) { std::fprintf(stderr, "[srla_b200] size pass failed: %s\n", cudaGetErrorString(cudaGetLastError())); return SRLA_APIRESULT_NG; }
                const JobOut *jo = (const JobOut *)c->h_jobout.p;
                for (uint32_t k = 0; k < cnt; k++) { if (jo[k].status) { return SRLA_APIRESULT_NG; } est[b + k] = jo[k].estimate_bytes; }
            }
            /* shortest path per chunk -> final block list */
            jobs.clear();
            for (Chunk &ch : chunks) {
                std::vector<uint32_t> edge((size_t)ch.nodes * ch.nodes, 0u);
                uint32_t k = ch.first_job;
                for (uint32_t i = 0; i < ch.nodes; i++) {
                    for (uint32_t j = i + 1; j < ch.nodes; j++) {
                        if ((j - i) * min_block > max_block) { continue; }
                        edge[i * ch.nodes + j] = est[k++];
                    }
                }
                std::vector<uint32_t> parts;
                shortest_partition(edge, ch.nodes, min_block, ch.len, parts);
                uint32_t off = 0;
                ch.final_begin = (uint32_t)jobs.size();
                for (uint32_t len : parts) {
                    jobs.push_back(make_job(ch.stream, ch.at + off, len, (ch.at + off == 0) ? kJobFirstOfStream : 0u, len_cache));
                    off += len;
                }
                ch.final_end = (uint32_t)jobs.size();
            }
            /* the coded blocks that depend on earlier calls: the call list in front of such a block is the previous chunk's
             * coded blocks, this chunk's whole search, and this chunk's coded blocks in front of it */
            c->tails_scratch.clear();
            std::vector<Job> call_list;
            if (var_tails) {
                for (size_t ci = 0; ci < chunks.size(); ci++) {
                    const Chunk &ch = chunks[ci];
                    bool any = false;
                    for (uint32_t k = ch.final_begin; k < ch.final_end; k++) { any = any || dependent(jobs[k].nsmpl); }
                    if (!any) { continue; }
                    const uint32_t first = (uint32_t)call_list.size();
                    if (ci > 0 && chunks[ci - 1].stream == ch.stream) {
                        for (uint32_t k = chunks[ci - 1].final_begin; k < chunks[ci - 1].final_end; k++) { call_list.push_back(jobs[k]); }
                    }
                    for (uint32_t k = 0; k < ch.num_jobs; k++) { call_list.push_back(cand_jobs[ch.first_job + k]); }
                    for (uint32_t k = ch.final_begin; k < ch.final_end; k++) {
                        if (dependent(jobs[k].nsmpl)) {
                            TailJob tj; tj.job = (uint32_t)call_list.size(); tj.first_of_stream = first; tj.out = k;
                            c->tails_scratch.push_back(tj);
                        }
                        call_list.push_back(jobs[k]);
                    }
                }
            }
            if (!upload_jobs(jobs)) { return SRLA_APIRESULT_NG; }
            if (!c->tails_scratch.empty()) {
                const size_t bytes = sizeof(Job) * call_list.size();
                if (!c->chain.reserve(bytes)) { return SRLA_APIRESULT_NG; }
                if (cudaMemcpyAsync(c->chain.p, call_list.data(), bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess
                    || cudaStreamSynchronize(c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }       /* call_list is a local */
                if (!upload_tails()) { return SRLA_APIRESULT_NG; }
                var_final_tails = true;
            }
        }
        stt.num_analysed += jobs.size();
        stt.num_blocks = pl.size_only ? 0 : jobs.size();
        for (const Job &j : jobs) { nmax = std::max(nmax, j.nsmpl); }
        if (var_final_tails) { nmax = max_block; }            /* the replayed search segments can be longer than any coded block */

        /* ---- groups: [lshift so far] -> analyse -> decide -> scan -> emit ---- */
        if (pl.variable) {                      /* the final block list replaces the fixed tiling */
            gstart.clear();
            for (size_t at = 0; at < jobs.size(); at += per_batch) { gstart.push_back(at); }
            gstart.push_back(jobs.size());
        }
        const size_t groups_now = gstart.size() - 1;
        std::vector<cudaEvent_t> &grp_done = c->ev_grp;
        unsigned long long *mailbox = nullptr; uint32_t *d_snap = nullptr;
        if (pipelined) {
            while (grp_done.size() < groups_now) { cudaEvent_t e; if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return SRLA_APIRESULT_NG; } grp_done.push_back(e); }
            if (!c->h_mailbox.reserve(sizeof(unsigned long long) * (groups_now + 1))) { return SRLA_APIRESULT_NG; }
            mailbox = (unsigned long long *)c->h_mailbox.p;
            if (!c->snapshot.reserve(sizeof(uint32_t) * (groups_now + 1) * pl.num_streams)) { return SRLA_APIRESULT_NG; }
            d_snap = (uint32_t *)c->snapshot.p;
        }
        /* device -> host copies of finished groups, in order.  With the feeder the copies land in pinned staging and
         * its threads move the bytes on to the caller's (pageable) buffer as they arrive; otherwise they go straight to
         * the caller's buffer.  drain(false) is polled while the host queues groups or waits for the feeder, so a
         * finished group's bytes leave as soon as they exist. */
        bool host_overflow = false, drain_failed = false;
        const bool staged = pipelined && io->narrow && feeder.out_stage != nullptr;
        uint8_t *h_dst = pipelined ? (staged ? (uint8_t *)c->h_stage_out.p : io->out) : nullptr;
        if (staged) { while (c->ev_d2h.size() < groups_now) { cudaEvent_t e; if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return SRLA_APIRESULT_NG; } c->ev_d2h.push_back(e); } }
        std::vector<unsigned long long> gend(groups_now, 0);
        size_t published = 0, drained = 0, recorded = 0;
        unsigned long long prev = 0;
        auto drain = [&](bool block) {
            while (drained < recorded && !host_overflow && !drain_failed) {
                if (block) {
                    if (cudaEventSynchronize(grp_done[drained]) != cudaSuccess) { drain_failed = true; break; }
                } else {
                    const cudaError_t q = cudaEventQuery(grp_done[drained]);
                    if (q == cudaErrorNotReady) { (void)cudaGetLastError(); break; }
                    if (q != cudaSuccess) { drain_failed = true; break; }
                }
                const unsigned long long end = mailbox[drained];
                if (end > io->out_capacity || end > cap) { host_overflow = true; break; }
                if (end > prev && cudaMemcpyAsync(h_dst + prev, d_out + prev, end - prev, cudaMemcpyDeviceToHost, c->d2h_stream) != cudaSuccess) { drain_failed = true; break; }
                prev = end; gend[drained] = end;
                if (c->trace) { host_d2h.push_back(host_ms()); }
                if (staged && cudaEventRecord(c->ev_d2h[drained], c->d2h_stream) != cudaSuccess) { drain_failed = true; break; }
                drained++;
            }
            if (staged) {
                const size_t before = published;
                while (published < drained && cudaEventQuery(c->ev_d2h[published]) == cudaSuccess) { feeder.out_ready.store(gend[published], std::memory_order_release); published++; }
                (void)cudaGetLastError();                      /* cudaEventQuery's cudaErrorNotReady is not an error */
                if (on_ready && published > before) { on_ready(h_dst, gend[published - 1]); }
            }
        };
        if (pipelined) { poll_d2h = [&] { drain(false); }; }
        const int lanes_now = pl.variable ? 1 : lanes;
        if (lanes_now > 1) {
            /* fork: the other lanes start after everything queued on the caller's stream so far (uploads, shift) */
            if (cudaEventRecord(c->ev_fork, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
            for (int l = 1; l < lanes_now; l++) { if (cudaStreamWaitEvent(c->lane[l].stream, c->ev_fork, 0) != cudaSuccess) { return SRLA_APIRESULT_NG; } }
            while (c->ev_scan.size() < groups_now) { cudaEvent_t e; if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return SRLA_APIRESULT_NG; } c->ev_scan.push_back(e); }
        }
        for (size_t g = 0; g < groups_now; g++) {
            const size_t j0 = gstart[g];
            const uint32_t cnt = (uint32_t)(gstart[g + 1] - j0);
            const int ln = (int)(g % (size_t)lanes_now);
            const cudaStream_t on = c->lane[ln].stream;
            if (pipelined) {
                /* pinned source: queue the next group's copy first so the copy engine never waits for the host */
                if (copies_async && !io->narrow && g + 1 < groups_now && !issue_h2d(g + 1)) { return SRLA_APIRESULT_NG; }
                if (cudaStreamWaitEvent(on, h2d_done[g], 0) != cudaSuccess) { return SRLA_APIRESULT_NG; }
                if (!launch_lshift(pl, (const Job *)c->jobs.p + j0, cnt, d_snap + g * pl.num_streams, on, ingest)) { return SRLA_APIRESULT_NG; }
            }
            const bool chain = lanes_now > 1;
            if (!run_batch(pl, (const Job *)c->jobs.p + j0, cnt, nmax, !pl.size_only, d_out, cap, !pl.size_only, ev_idx++,
                           pipelined ? mailbox + g : nullptr, ln,
                           (chain && g > 0) ? c->ev_scan[g - 1] : nullptr, chain ? c->ev_scan[g] : nullptr, use_tails || var_final_tails,
                           var_final_tails ? (const Job *)c->chain.p : nullptr)) { return SRLA_APIRESULT_NG; }
            if (pipelined && cudaEventRecord(grp_done[g], on) != cudaSuccess) { return SRLA_APIRESULT_NG; }
            if (c->trace) { host_launch.push_back(host_ms()); }
            if (pipelined) { recorded = g + 1; drain(false); }
            /* pageable source (the copy blocks the host) or feeder staging: the next group's copy follows this
             * group's launches, so the device works on group g while the host moves group g + 1 */
            if (pipelined && (!copies_async || io->narrow) && g + 1 < groups_now && !issue_h2d(g + 1)) { return SRLA_APIRESULT_NG; }
        }
        /* join: the caller's stream continues after every lane */
        for (int l = 1; l < lanes_now; l++) {
            if (cudaEventRecord(c->lane[l].done, c->lane[l].stream) != cudaSuccess || cudaStreamWaitEvent(c->stream, c->lane[l].done, 0) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        }
        if (pipelined) {
            /* the shift every stream really has, once all groups have been OR-ed in (slot groups_now) */
            lshift_finish_kernel<<<(pl.num_streams + 255) / 256, 256, 0, c->stream>>>((StreamDev *)c->streams.p, pl.num_streams, d_snap + groups_now * pl.num_streams);
            launches++;
        }
        if (cudaEventRecord(c->ev_end, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }

        /* ---- the remaining device -> host copies ---- */
        poll_d2h = nullptr;
        if (pipelined) {
            drain(true);
            if (drain_failed) { std::fprintf(stderr, "[srla_b200] encode failed on the device: %s\n", cudaGetErrorString(cudaGetLastError())); return SRLA_APIRESULT_NG; }
            if (staged) {
                if (host_overflow) { feeder.abort.store(1, std::memory_order_release); }
                else {
                    for (; published < groups_now; published++) {
                        if (cudaEventSynchronize(c->ev_d2h[published]) != cudaSuccess) { return SRLA_APIRESULT_NG; }
                        feeder.out_ready.store(gend[published], std::memory_order_release);
                    }
                    feeder.out_total.store(prev, std::memory_order_release);
                }
                feeder.join();
            }
        }

        /* ---- results ---- */
        const size_t small_bytes = kMiscBytes;
        const size_t sb_bytes = sizeof(unsigned long long) * (pl.num_streams + 1);
        const size_t snap_bytes = pipelined ? sizeof(uint32_t) * (groups_now + 1) * pl.num_streams : 0;
        if (!c->h_result.reserve(small_bytes + sb_bytes + sizeof(JobOut) + snap_bytes + 64)) { return SRLA_APIRESULT_NG; }
        unsigned char *hs = (unsigned char *)c->h_result.p;
        cudaMemcpyAsync(hs, c->misc.p, small_bytes + sb_bytes, cudaMemcpyDeviceToHost, c->stream);      /* counters, statistics, stream offsets */
        if (single_estimate) { cudaMemcpyAsync(hs + small_bytes + sb_bytes, c->lane[0].jobout.p, sizeof(JobOut), cudaMemcpyDeviceToHost, c->stream); }
        if (pipelined) { cudaMemcpyAsync(hs + small_bytes + sb_bytes + sizeof(JobOut), c->snapshot.p, snap_bytes, cudaMemcpyDeviceToHost, c->stream); }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess || (pipelined && (cudaStreamSynchronize(c->d2h_stream) != cudaSuccess || cudaStreamSynchronize(c->copy_stream) != cudaSuccess))) {
            std::fprintf(stderr, "[srla_b200] encode failed on the device: %s\n", cudaGetErrorString(cudaGetLastError()));
            return SRLA_APIRESULT_NG;
        }
        if (io && io->narrow && feeder.overflow.load()) { narrow_overflow = true; return SRLA_APIRESULT_NG; }
        const unsigned long long *running = (const unsigned long long *)hs;
        const uint32_t *dstats = (const uint32_t *)(hs + 2 * sizeof(unsigned long long));
        const unsigned long long *sbeg = (const unsigned long long *)(hs + small_bytes);
        const JobOut *first = (const JobOut *)(hs + small_bytes + sb_bytes);
        if (pipelined) {
            /* did every group run with the final shift of the streams it touched? */
            const uint32_t *snap = (const uint32_t *)(hs + small_bytes + sb_bytes + sizeof(JobOut));
            const uint32_t *fin = snap + groups_now * pl.num_streams;
            bool redo = false;
            for (size_t g = 0; g < groups_now && !redo; g++) {
                const size_t j0 = gstart[g], j1 = gstart[g + 1];
                for (size_t j = j0; j < j1; j++) { const uint32_t s = jobs[j].stream; if (snap[g * pl.num_streams + s] != fin[s]) { redo = true; break; } }
            }
            if (redo) {
                Plan again = pl; again.allow_pipeline = false;
                Runner r2{ enc, c };
                const SRLAApiResult rc = r2.run(again, d_out, cap, stream_offsets, single_estimate, nullptr);
                if (rc != SRLA_APIRESULT_OK) { return rc; }
                const uint64_t total = enc->stats.bytes_out;
                if (total > io->out_capacity) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
                if (cudaMemcpy(io->out, d_out, total, cudaMemcpyDeviceToHost) != cudaSuccess) { return SRLA_APIRESULT_NG; }
                return SRLA_APIRESULT_OK;
            }
        }
        stt.kernel_launches = launches;
        stt.bytes_out = running[0];
        for (int i = 0; i < 256; i++) { stt.order_histogram[i] = dstats[i]; }
        for (int i = 0; i < 4; i++) { stt.method_histogram[i] = dstats[256 + i]; }
        for (int i = 0; i < 3; i++) { stt.type_histogram[i] = dstats[260 + i]; }
        cudaEventElapsedTime(&stt.ms_total_device, c->ev_begin, c->ev_end);
        if (c->trace && pipelined) {
            std::fprintf(stderr, "[srla_b200 trace] %zu groups, %d lanes, total %.3f ms (device clock, origin = call start); host clock at return %.3f ms\n", groups_now, lanes_now, stt.ms_total_device, host_ms());
            for (size_t g = 0; g < groups_now; g++) {
                float h = 0, t[5] = { 0, 0, 0, 0, 0 };
                cudaEventElapsedTime(&h, c->ev_begin, h2d_done[g]);
                for (int k = 0; k < 5; k++) { cudaEventElapsedTime(&t[k], c->ev_begin, c->ev_pool[g * 5 + k]); }
                std::fprintf(stderr, "  group %2zu jobs %5zu lane %d: h2d done %.3f | front %.3f lpc %.3f residual %.3f decide %.3f end %.3f | host: h2d issued %.3f, kernels issued %.3f, d2h issued %.3f\n",
                             g, gstart[g + 1] - gstart[g], (int)(g % (size_t)lanes_now), h, t[0], t[1], t[2], t[3], t[4],
                             g < host_h2d.size() ? host_h2d[g] : -1.0, g < host_launch.size() ? host_launch[g] : -1.0, g < host_d2h.size() ? host_d2h[g] : -1.0);
            }
        }
        for (size_t i = 0; i < ev_idx; i++) {
            float t[4] = { 0, 0, 0, 0 };
            for (int k = 0; k < 4; k++) { cudaEventElapsedTime(&t[k], c->ev_pool[i * 5 + k], c->ev_pool[i * 5 + k + 1]); }
            stt.ms_front += t[0]; stt.ms_lpc += t[1]; stt.ms_residual += t[2]; stt.ms_emit += t[3];
            stt.ms_analyse += t[0] + t[1] + t[2];
        }
        if (single_estimate) { *single_estimate = first->estimate_bytes; if (first->status) { return SRLA_APIRESULT_NG; } }
        if (pl.size_only) { return SRLA_APIRESULT_OK; }
        /* a block whose analysis failed the way the reference fails (singular LTP system) is never
         * written: every emitted job increments exactly one block-type counter */
        {
            const uint64_t emitted = (uint64_t)dstats[260] + dstats[261] + dstats[262];
            if (running[1] == 0 && !host_overflow && emitted != jobs.size()) { return SRLA_APIRESULT_NG; }
        }
        if (running[1] != 0 || running[0] > cap || host_overflow) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
        if (io) {
            if (running[0] > io->out_capacity) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
            if (!pipelined && cudaMemcpy(io->out, d_out, running[0], cudaMemcpyDeviceToHost) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        }
        if (stream_offsets) {
            for (uint32_t s = 0; s < pl.num_streams; s++) { stream_offsets[s] = pl.emit_stream_header ? sbeg[s] : 0; }
            stream_offsets[pl.num_streams] = running[0];
        }
        return SRLA_APIRESULT_OK;
    }
};

SRLAApiResult check_ready(const SRLAEncoder *e)
{
    if (!e || e->magic != kEncoderMagic) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (e->set_parameter != 1) { return SRLA_APIRESULT_PARAMETER_NOT_SET; }
    return SRLA_APIRESULT_OK;
}

uint64_t max_encoded_size(const SRLAEncoder *e, uint32_t num_samples)
{
    const uint32_t unit = (e->param.min_num_samples_per_block == e->param.max_num_samples_per_block)
        ? e->param.max_num_samples_per_block : e->param.min_num_samples_per_block;
    const uint64_t blocks = ((uint64_t)num_samples + unit - 1) / unit;
    return 30ull + blocks * 11ull + ((uint64_t)e->param.bits_per_sample * num_samples * e->param.num_channels + 7) / 8 + 64;
}

/* host PCM (planar pointers) -> device, as one StreamDev-style layout */
bool upload_planar_int32(DeviceCtx *c, const int32_t *const *input, uint32_t nch, uint32_t n, struct SRLAB200Stream *desc)
{
    const uint64_t stride = round_up_u32(n, 8);
    if (!c->pcm.reserve(sizeof(int32_t) * stride * nch)) { return false; }
    for (uint32_t ch = 0; ch < nch; ch++) {
        CU_TRY(cudaMemcpyAsync((int32_t *)c->pcm.p + stride * ch, input[ch], sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    }
    desc->pcm = c->pcm.p; desc->channel_stride = stride; desc->num_samples = n; desc->sample_bytes = 4;
    return true;
}

} // namespace

/* ================================================================================================
 * Part 1: reference-compatible surface
 * ============================================================================================== */
extern "C" {

SRLAApiResult SRLAEncoder_EncodeHeader(const struct SRLAHeader *header, uint8_t *data, uint32_t data_size)
{
    if (header == NULL || data == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (data_size < SRLA_HEADER_SIZE) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    if (header->num_channels == 0 || header->num_samples == 0 || header->sampling_rate == 0 || header->bits_per_sample == 0
        || header->offset_lshift >= 32 || header->max_num_samples_per_block == 0 || header->preset >= SRLA_NUM_PARAMETER_PRESETS) {
        return SRLA_APIRESULT_INVALID_FORMAT;
    }
    uint8_t *p = data;
    auto be = [&p](uint32_t v, int nbytes) { for (int i = nbytes - 1; i >= 0; i--) { *p++ = (uint8_t)(v >> (8 * i)); } };
    *p++ = '1'; *p++ = '2'; *p++ = '4'; *p++ = '9';
    be(SRLA_FORMAT_VERSION, 4); be(SRLA_CODEC_VERSION, 4);
    be(header->num_channels, 2); be(header->num_samples, 4); be(header->sampling_rate, 4); be(header->bits_per_sample, 2);
    be(header->offset_lshift, 1); be(header->max_num_samples_per_block, 4); be(header->preset, 1);
    return SRLA_APIRESULT_OK;
}

int32_t SRLAEncoder_CalculateWorkSize(const struct SRLAEncoderConfig *config)
{
    if (!config_valid(config)) { return -1; }
    return (int32_t)(sizeof(struct SRLAEncoder) + 64);
}

struct SRLAEncoder *SRLAEncoder_Create(const struct SRLAEncoderConfig *config, void *work, int32_t work_size)
{
    uint8_t own = 0;
    if (work == NULL && work_size == 0) {
        if ((work_size = SRLAEncoder_CalculateWorkSize(config)) < 0) { return NULL; }
        work = std::malloc((size_t)work_size);
        own = 1;
    }
    if (config == NULL || work == NULL || work_size < SRLAEncoder_CalculateWorkSize(config) || !config_valid(config)) {
        if (own) { std::free(work); }
        return NULL;
    }
    if (config->max_num_samples_per_block > (uint32_t)kMaxBlock) {                 /* the block header's sample count is 16 bits */
        std::fprintf(stderr, "[srla_b200] max_num_samples_per_block %u exceeds this implementation's capacity %d\n", config->max_num_samples_per_block, kMaxBlock);
        if (own) { std::free(work); }
        return NULL;
    }
    uintptr_t at = ((uintptr_t)work + 15u) & ~(uintptr_t)15u;
    struct SRLAEncoder *e = (struct SRLAEncoder *)at;
    std::memset(e, 0, sizeof(*e));
    e->magic = kEncoderMagic;
    e->config = *config;
    e->alloced_by_own = own;
    e->work = work;
    e->ctx = new (std::nothrow) DeviceCtx();
    if (!e->ctx || !ctx_init(e->ctx)) {
        if (e->ctx) { ctx_destroy(e->ctx); delete e->ctx; }
        e->magic = 0;
        if (own) { std::free(work); }
        return NULL;
    }
    return e;
}

void SRLAEncoder_Destroy(struct SRLAEncoder *encoder)
{
    if (encoder == NULL || encoder->magic != kEncoderMagic) { return; }
    ctx_destroy(encoder->ctx);
    delete encoder->ctx;
    encoder->ctx = nullptr;
    encoder->magic = 0;
    if (encoder->alloced_by_own == 1) { std::free(encoder->work); }
}

SRLAApiResult SRLAEncoder_SetEncodeParameter(struct SRLAEncoder *encoder, const struct SRLAEncodeParameter *parameter)
{
    if (encoder == NULL || parameter == NULL || encoder->magic != kEncoderMagic) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    /* srla_encoder.c:427-465 */
    if (parameter->num_channels == 0 || parameter->bits_per_sample == 0 || parameter->sampling_rate == 0
        || parameter->preset >= SRLA_NUM_PARAMETER_PRESETS) { return SRLA_APIRESULT_INVALID_FORMAT; }
    /* srla_encoder.c:727-734 */
    if (parameter->min_num_samples_per_block == 0
        || parameter->min_num_samples_per_block > parameter->max_num_samples_per_block
        || parameter->num_lookahead_samples < parameter->max_num_samples_per_block
        || (parameter->num_lookahead_samples % parameter->min_num_samples_per_block) != 0
        || (parameter->ltp_order > 0 && (parameter->ltp_order % 2) == 0) || parameter->ltp_order > SRLA_MAX_LTP_ORDER) {
        return SRLA_APIRESULT_INVALID_FORMAT;
    }
    /* this implementation: raw blocks exist for 8/16/24 bit only */
    if (parameter->bits_per_sample != 8 && parameter->bits_per_sample != 16 && parameter->bits_per_sample != 24) { return SRLA_APIRESULT_INVALID_FORMAT; }
    /* srla_encoder.c:737-742 */
    if (encoder->config.max_num_samples_per_block < parameter->max_num_samples_per_block
        || encoder->config.min_num_samples_per_block > parameter->min_num_samples_per_block
        || encoder->config.max_num_lookahead_samples < parameter->num_lookahead_samples
        || encoder->config.max_num_channels < parameter->num_channels
        || parameter->num_channels > SRLA_MAX_NUM_CHANNELS
        || encoder->config.max_num_parameters < kPresetMaxOrder[parameter->preset]) {
        return SRLA_APIRESULT_INSUFFICIENT_BUFFER;
    }
    encoder->param = *parameter;
    encoder->max_order = kPresetMaxOrder[parameter->preset];
    encoder->offset_lshift = 0;
    encoder->set_parameter = 1;
    return SRLA_APIRESULT_OK;
}

static SRLAApiResult single_chunk(struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
                                  uint8_t *data, uint32_t data_size, uint32_t *output_size, bool variable, bool size_only)
{
    DeviceCtx *c = encoder->ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    struct SRLAB200Stream desc;
    if (!upload_planar_int32(c, input, encoder->param.num_channels, num_samples, &desc)) { return SRLA_APIRESULT_NG; }
    Plan pl;
    pl.streams = &desc; pl.num_streams = 1; pl.nch = encoder->param.num_channels;
    pl.emit_stream_header = false; pl.use_fixed_lshift = true; pl.fixed_lshift = encoder->offset_lshift;
    pl.variable = variable; pl.size_only = size_only;
    Runner r{ encoder, c };
    const uint64_t cap = max_encoded_size(encoder, num_samples);
    if (!size_only && !c->out.reserve(cap)) { return SRLA_APIRESULT_NG; }
    uint64_t offs[2] = { 0, 0 };
    uint32_t est = 0;
    const SRLAApiResult rc = r.run(pl, size_only ? nullptr : (uint8_t *)c->out.p, size_only ? 0 : cap, offs, size_only ? &est : nullptr, nullptr);
    if (rc != SRLA_APIRESULT_OK) { return rc; }
    if (size_only) { *output_size = est; return SRLA_APIRESULT_OK; }
    if (offs[1] > data_size) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    if (cudaMemcpy(data, c->out.p, offs[1], cudaMemcpyDeviceToHost) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    *output_size = (uint32_t)offs[1];
    return SRLA_APIRESULT_OK;
}

SRLAApiResult SRLAEncoder_ComputeBlockSize(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples, uint32_t *output_size)
{
    if (encoder == NULL || input == NULL || num_samples == 0 || output_size == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    if (num_samples > encoder->param.max_num_samples_per_block) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    return single_chunk(encoder, input, num_samples, nullptr, 0, output_size, false, true);
}

SRLAApiResult SRLAEncoder_EncodeBlock(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
    uint8_t *data, uint32_t data_size, uint32_t *output_size)
{
    if (encoder == NULL || input == NULL || num_samples == 0 || data == NULL || data_size == 0 || output_size == NULL) {
        return SRLA_APIRESULT_INVALID_ARGUMENT;
    }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    if (num_samples > encoder->param.max_num_samples_per_block) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    return single_chunk(encoder, input, num_samples, data, data_size, output_size, false, false);
}

SRLAApiResult SRLAEncoder_EncodeOptimalPartitionedBlock(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
    uint8_t *data, uint32_t data_size, uint32_t *output_size)
{
    if (encoder == NULL || input == NULL || data == NULL || output_size == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    if (num_samples == 0 || num_samples > encoder->param.num_lookahead_samples) { return SRLA_APIRESULT_NG; }
    /* one look-ahead chunk: force a single chunk by planning with step == num_lookahead_samples */
    return single_chunk(encoder, input, num_samples, data, data_size, output_size, true, false);
}

SRLAApiResult SRLAEncoder_EncodeWhole(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
    uint8_t *data, uint32_t data_size, uint32_t *output_size, SRLAEncoder_EncodeBlockCallback encode_callback)
{
    if (encoder == NULL || input == NULL || data == NULL || output_size == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    if (num_samples == 0) { return SRLA_APIRESULT_INVALID_FORMAT; }
    if (data_size < SRLA_HEADER_SIZE) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    DeviceCtx *c = encoder->ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    const uint32_t nch = encoder->param.num_channels;
    const uint64_t stride = round_up_u32(num_samples, 16);
    const uint64_t cap = max_encoded_size(encoder, num_samples);
    if (!c->pcm.reserve(sizeof(int32_t) * stride * nch) || !c->out.reserve(cap)) { return SRLA_APIRESULT_NG; }
    HostStream hs;
    for (uint32_t ch = 0; ch < nch; ch++) { if (input[ch] == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; } hs.ch[ch] = input[ch]; }
    Plan pl;
    pl.num_streams = 1; pl.nch = nch;
    pl.variable = encoder->param.min_num_samples_per_block != encoder->param.max_num_samples_per_block;
    /* long fixed-block inputs of <= 16-bit sources are narrowed to int16 by the host feeder on their way in */
    const uint64_t num_blocks = ((uint64_t)num_samples + encoder->param.max_num_samples_per_block - 1) / encoder->param.max_num_samples_per_block;
    bool narrow = c->feed_threads > 0 && encoder->param.bits_per_sample <= 16 && !pl.variable && num_blocks >= 2048
                  && c->h_stage.reserve(sizeof(int16_t) * stride * nch);
    uint64_t offs[2] = { 0, 0 };
    SRLAApiResult rc = SRLA_APIRESULT_NG;
    /* progress callbacks, one per top-level step (srla_encoder.c:1756-1783): `report` walks the block headers of the bytes
     * that exist so far and calls back for every step that is complete.  On the pipelined path it runs while later groups
     * are still being encoded (a progress display moves during a long file); whatever is left is reported after the call. */
    const uint32_t cb_step = pl.variable ? encoder->param.num_lookahead_samples : encoder->param.max_num_samples_per_block;
    uint32_t cb_progress = 0; uint64_t cb_pos = SRLA_HEADER_SIZE;
    auto report = [&](const uint8_t *bytes, uint64_t ready) {
        if (encode_callback == NULL) { return; }
        while (cb_progress < num_samples) {
            const uint32_t todo = std::min(cb_step, num_samples - cb_progress);
            uint32_t got = 0; uint64_t pos = cb_pos;
            while (got < todo && pos + 11 <= ready) {
                const uint32_t size = ((uint32_t)bytes[pos + 2] << 24) | ((uint32_t)bytes[pos + 3] << 16) | ((uint32_t)bytes[pos + 4] << 8) | bytes[pos + 5];
                if (pos + 6ull + size > ready) { return; }                    /* the block is not complete yet */
                got += ((uint32_t)bytes[pos + 9] << 8) | bytes[pos + 10];
                pos += 6ull + size;
            }
            if (got < todo) { return; }
            cb_progress += todo;
            encode_callback(num_samples, cb_progress, bytes + cb_pos, (uint32_t)(pos - cb_pos));
            cb_pos = pos;
        }
    };
    for (;;) {
        struct SRLAB200Stream desc;
        desc.pcm = c->pcm.p; desc.channel_stride = stride; desc.num_samples = num_samples; desc.sample_bytes = narrow ? 2u : 4u;
        HostIO io; io.streams = &hs; io.out = data; io.out_capacity = data_size; io.narrow = narrow;
        pl.streams = &desc;
        Runner r{ encoder, c };
        if (encode_callback != NULL) { r.on_ready = report; }
        rc = r.run(pl, (uint8_t *)c->out.p, cap, offs, nullptr, &io);
        if (narrow && r.narrow_overflow) { narrow = false; continue; }       /* samples beyond 16 bits: int32 layout */
        break;
    }
    if (rc != SRLA_APIRESULT_OK) { return rc; }
    *output_size = (uint32_t)offs[1];
    encoder->offset_lshift = data[24];                        /* the reference keeps it in its header (srla_encoder.c:1732) */
    report(data, offs[1]);                                    /* the steps not reported while the call ran */
    return SRLA_APIRESULT_OK;
}

/* ================================================================================================
 * Part 2: batch / device-resident extension
 * ============================================================================================== */
SRLAApiResult SRLAB200_EncodeStreamsDevice(
    struct SRLAEncoder *encoder, const struct SRLAB200Stream *streams, uint32_t num_streams,
    uint8_t *d_out, uint64_t out_capacity, uint64_t *stream_offsets)
{
    if (encoder == NULL || streams == NULL || num_streams == 0 || d_out == NULL || stream_offsets == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    Plan pl;
    pl.streams = streams; pl.num_streams = num_streams; pl.nch = encoder->param.num_channels;
    pl.variable = encoder->param.min_num_samples_per_block != encoder->param.max_num_samples_per_block;
    Runner r{ encoder, encoder->ctx };
    return r.run(pl, d_out, out_capacity, stream_offsets, nullptr, nullptr);
}

SRLAApiResult SRLAB200_EncodeStreamsHost(
    struct SRLAEncoder *encoder, const struct SRLAB200Stream *streams, uint32_t num_streams,
    uint8_t *out, uint64_t out_capacity, uint64_t *stream_offsets)
{
    if (encoder == NULL || streams == NULL || num_streams == 0 || out == NULL || stream_offsets == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    DeviceCtx *c = encoder->ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    const uint32_t nch = encoder->param.num_channels;
    std::vector<struct SRLAB200Stream> dev(num_streams);
    std::vector<HostStream> host(num_streams);
    uint64_t total_bytes = 0, cap = 0;
    for (uint32_t s = 0; s < num_streams; s++) {
        if (streams[s].pcm == NULL || (streams[s].sample_bytes != 2 && streams[s].sample_bytes != 4)) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
        total_bytes += (uint64_t)round_up_u32(streams[s].num_samples, 16) * nch * streams[s].sample_bytes;
        cap += max_encoded_size(encoder, streams[s].num_samples);
    }
    if (!c->pcm.reserve(total_bytes) || !c->out.reserve(cap)) { return SRLA_APIRESULT_NG; }
    uint64_t at = 0;
    for (uint32_t s = 0; s < num_streams; s++) {
        const uint64_t stride = round_up_u32(streams[s].num_samples, 16);
        const uint32_t sb = streams[s].sample_bytes;
        dev[s].pcm = (unsigned char *)c->pcm.p + at; dev[s].channel_stride = stride;
        dev[s].num_samples = streams[s].num_samples; dev[s].sample_bytes = sb;
        for (uint32_t ch = 0; ch < nch; ch++) { host[s].ch[ch] = (const unsigned char *)streams[s].pcm + streams[s].channel_stride * ch * sb; }
        at += stride * nch * sb;
    }
    HostIO io; io.streams = host.data(); io.out = out; io.out_capacity = out_capacity;
    Plan pl;
    pl.streams = dev.data(); pl.num_streams = num_streams; pl.nch = nch;
    pl.variable = encoder->param.min_num_samples_per_block != encoder->param.max_num_samples_per_block;
    Runner r{ encoder, c };
    return r.run(pl, (uint8_t *)c->out.p, cap, stream_offsets, nullptr, &io);
}

SRLAApiResult SRLAB200_EncodeInterleavedHost(
    struct SRLAEncoder *encoder, const struct SRLAB200Frames *items, uint32_t num_streams,
    uint8_t *out, uint64_t out_capacity, uint64_t *stream_offsets)
{
    if (encoder == NULL || items == NULL || num_streams == 0 || out == NULL || stream_offsets == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    DeviceCtx *c = encoder->ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    const uint32_t nch = encoder->param.num_channels;
    const uint32_t cb = encoder->param.bits_per_sample / 8u;              /* 1, 2 or 3: SetEncodeParameter admits 8/16/24 bits */
    const uint32_t sb = (encoder->param.bits_per_sample <= 16) ? 2u : 4u; /* planar device layout */
    std::vector<struct SRLAB200Stream> dev(num_streams);
    std::vector<const void *> raw_host(num_streams), raw_dev(num_streams);
    uint64_t planar_bytes = 0, raw_bytes = 0, cap = 0;
    for (uint32_t s = 0; s < num_streams; s++) {
        if (items[s].frames == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
        planar_bytes += (uint64_t)round_up_u32(items[s].num_samples, 16) * nch * sb;
        raw_bytes += ((uint64_t)items[s].num_samples * nch * cb + 255u) / 256u * 256u;
        cap += max_encoded_size(encoder, items[s].num_samples);
    }
    if (!c->pcm.reserve(planar_bytes) || !c->raw.reserve(raw_bytes) || !c->out.reserve(cap)) { return SRLA_APIRESULT_NG; }
    uint64_t at = 0, raw_at = 0;
    for (uint32_t s = 0; s < num_streams; s++) {
        const uint64_t stride = round_up_u32(items[s].num_samples, 16);
        dev[s].pcm = (unsigned char *)c->pcm.p + at; dev[s].channel_stride = stride;
        dev[s].num_samples = items[s].num_samples; dev[s].sample_bytes = sb;
        raw_host[s] = items[s].frames; raw_dev[s] = (unsigned char *)c->raw.p + raw_at;
        at += stride * nch * sb;
        raw_at += ((uint64_t)items[s].num_samples * nch * cb + 255u) / 256u * 256u;
    }
    HostIO io; io.raw = raw_host.data(); io.out = out; io.out_capacity = out_capacity;
    Plan pl;
    pl.streams = dev.data(); pl.num_streams = num_streams; pl.nch = nch;
    pl.raw_dev = raw_dev.data(); pl.container_bytes = cb;
    pl.variable = encoder->param.min_num_samples_per_block != encoder->param.max_num_samples_per_block;
    Runner r{ encoder, c };
    return r.run(pl, (uint8_t *)c->out.p, cap, stream_offsets, nullptr, &io);
}

void *SRLAB200_AllocPinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}

void SRLAB200_FreePinned(void *p) { if (p) { cudaFreeHost(p); } }

uint64_t SRLAB200_MaxEncodedSize(const struct SRLAEncoder *encoder, uint32_t num_samples)
{
    if (encoder == NULL || encoder->magic != kEncoderMagic || encoder->set_parameter != 1) { return 0; }
    return max_encoded_size(encoder, num_samples);
}

SRLAApiResult SRLAB200_GetStats(const struct SRLAEncoder *encoder, struct SRLAB200Stats *stats)
{
    if (encoder == NULL || stats == NULL || encoder->magic != kEncoderMagic) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    *stats = encoder->stats;
    return SRLA_APIRESULT_OK;
}

SRLAApiResult SRLAB200_SetDevice(int device_ordinal)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device_ordinal < 0 || device_ordinal >= count) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    g_device = device_ordinal;
    return SRLA_APIRESULT_OK;
}

int SRLAB200_GetDeviceCount(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return count;
}

SRLAApiResult SRLAB200_GetDevicePciBusId(int device_ordinal, char *buffer, int buffer_size)
{
    if (buffer == NULL || buffer_size < 16) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (device_ordinal < 0) { if (cudaGetDevice(&device_ordinal) != cudaSuccess) { return SRLA_APIRESULT_NG; } }
    if (cudaDeviceGetPCIBusId(buffer, buffer_size, device_ordinal) != cudaSuccess) { (void)cudaGetLastError(); return SRLA_APIRESULT_INVALID_ARGUMENT; }
    return SRLA_APIRESULT_OK;
}

SRLAApiResult SRLAB200_SetStream(struct SRLAEncoder *encoder, void *cuda_stream)
{
    if (encoder == NULL || encoder->magic != kEncoderMagic) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    encoder->ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : encoder->ctx->own_stream;
    encoder->ctx->lane[0].stream = encoder->ctx->stream;
    return SRLA_APIRESULT_OK;
}

/* host-only: the feeder's int32 -> int16 narrowing (AVX2 with streaming stores where available); nonzero when a
 * sample does not fit.  Exported so the CPU test suite can check it without a device. */
uint32_t SRLAB200_TestNarrow(const int32_t *src, int16_t *dst, uint32_t count)
{
    if (src == NULL || dst == NULL) { return 1u; }
    return narrow_chunk(src, dst, count);
}

const char *SRLAB200_Version(void) { return "srla_b200 0.1 sm_100a (format 10 / codec 18)"; }

SRLAApiResult SRLAB200_TestAnalyseChannel(
    struct SRLAEncoder *encoder, const int32_t *sig, uint32_t n, int32_t *residual, struct SRLAB200ChannelResult *result)
{
    if (encoder == NULL || sig == NULL || n == 0 || residual == NULL || result == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    const SRLAApiResult ready = check_ready(encoder);
    if (ready != SRLA_APIRESULT_OK) { return ready; }
    if (n > encoder->param.max_num_samples_per_block || n <= encoder->max_order) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    DeviceCtx *c = encoder->ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    struct SRLAB200Stream desc;
    const int32_t *rows[1] = { sig };
    if (!upload_planar_int32(c, rows, 1, n, &desc)) { return SRLA_APIRESULT_NG; }
    Plan pl;
    pl.streams = &desc; pl.num_streams = 1; pl.nch = 1; pl.emit_stream_header = false; pl.use_fixed_lshift = true; pl.fixed_lshift = 0;
    pl.want_diag = true;
    Runner r{ encoder, c };
    std::memset(&encoder->stats, 0, sizeof(encoder->stats));
    if (!r.prepare_streams(pl)) { return SRLA_APIRESULT_NG; }
    Job job = make_job(0, 0, n, 0, r.len_cache);
    c->jobs_cached = false;
    if (!c->jobs.reserve(sizeof(Job))) { return SRLA_APIRESULT_NG; }
    if (cudaMemcpyAsync(c->jobs.p, &job, sizeof(Job), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    /* an odd length (or, with LTP, fewer than 263 samples) is analysed the way a freshly created reference handle would */
    c->tails_scratch.clear();
    const uint32_t ltp = encoder->param.ltp_order;
    const bool tail = (n & 1u) || (ltp > 0u && ceil_pow2_host(n) < 263u);
    if (tail) {
        TailJob tj; tj.job = 0; tj.first_of_stream = 0; tj.out = 0;
        c->tails_scratch.push_back(tj);
        if (!c->tails.reserve(sizeof(TailJob)) || cudaMemcpyAsync(c->tails.p, c->tails_scratch.data(), sizeof(TailJob), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    }
    if (!r.run_batch(pl, (const Job *)c->jobs.p, 1, n, false, nullptr, 0, true, 0, nullptr, 0, nullptr, nullptr, tail)) { return SRLA_APIRESULT_NG; }
    CandOut co; CandDiag dg;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess
        || cudaMemcpy(&co, c->lane[0].cand.p, sizeof(co), cudaMemcpyDeviceToHost) != cudaSuccess
        || cudaMemcpy(&dg, c->lane[0].diag.p, sizeof(dg), cudaMemcpyDeviceToHost) != cudaSuccess
        || cudaMemcpy(residual, c->lane[0].residual.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess) {
        std::fprintf(stderr, "[srla_b200] TestAnalyseChannel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return SRLA_APIRESULT_NG;
    }
    if (co.status) { return SRLA_APIRESULT_NG; }
    std::memset(result, 0, sizeof(*result));
    result->pre_coef = co.pre_coef; result->pre_prev = co.pre_prev;
    result->order = co.order; result->rshift = co.rshift; result->use_sum = co.use_sum;
    for (uint32_t i = 0; i < co.order && i < SRLA_MAX_COEFFICIENT_ORDER; i++) { result->coef[i] = co.coef[i]; }
    result->ltp_period = co.ltp_period;
    for (int i = 0; i < 3; i++) { result->ltp_coef[i] = co.ltp_coef[i]; }
    result->code_type = co.code_type; result->porder = co.porder; result->residual_bits = co.residual_bits; result->total_bits = co.total_bits;
    for (uint32_t i = 0; i <= encoder->max_order; i++) { result->autocorr[i] = dg.autocorr[i]; result->error_vars[i] = dg.error_vars[i]; }
    for (uint32_t i = 0; i < co.order; i++) { result->lpc_double[i] = dg.lpc_double[i]; }
    return SRLA_APIRESULT_OK;
}

} /* extern "C" */

#include "decoder.cuh"
