/*
 * decoder.cuh -- the decode path: SRLADecoder_* with the reference's signatures
 * (include/srla_decoder.h:8-56) on the GPU.  Included by libsrla_b200.cu (one translation unit).
 *
 * A stream decodes block by block and nothing crosses a block boundary (srla_decoder.c:633-799), so
 * a whole stream is two launches over all of its blocks.
 *   parse     (decode_parse_kernel) the bitstream of a block is serial (the second channel starts where
 *             the first one's codes end): one thread per block walks it, the lanes of a warp walk
 *             different blocks in lockstep (one flat loop, one code per trip, branch-free refill) -- side
 *             information to a per-channel record, residual codes (srla_coder.c:596-690) straight into
 *             the output buffer.
 *   synthesis (decode_blocks_kernel) one CTA per block, one warp per channel.  Warp 0 sums the Fletcher-16
 *             checksum (srla_utility.c:36-60) in parallel and judges the header.  LPC synthesis
 *             (srla_lpc_synthesize.c:238-262) is a recurrence over samples; warp w runs channel w as a
 *             systolic filter: lane L owns the outputs m = L (mod 32) and keeps their partial sums, every
 *             finished sample is broadcast with one shuffle and each lane adds its tap's product -- one
 *             shuffle and one multiply-add on the dependent chain per sample instead of `order` of them.
 *             Long-term synthesis (:264-327) advances in chunks shorter than the pitch lag, de-emphasis
 *             (srla_utility.c:361-378) is a first-order recurrence left to one lane.
 *   finish    mid/side -> left/right (srla_utility.c:106-174) and the offset shift, all threads.
 * All integer arithmetic wraps like the reference's int32 code does on x86.
 */
#ifndef SRLA_B200_DECODER_CUH
#define SRLA_B200_DECODER_CUH

namespace srla {

struct DecBlock {
    unsigned long long offset;     /* byte offset of the block (its sync code) inside the data buffer */
    uint32_t bytes;                /* 6 + size field                                                   */
    uint32_t sample_offset;        /* where its samples go inside every channel                        */
    uint32_t nsmpl;                /* samples per channel (block header, read by the host walk)        */
    uint32_t pad;
};

struct DecParams {
    const uint8_t *data;
    const DecBlock *blocks;
    uint32_t *status;              /* per block: SRLAApiResult                                         */
    int32_t *out;                  /* planar: channel c at out + c * stride                            */
    unsigned long long stride;
    uint32_t nch, bps, lshift, check;
    const uint16_t *tree;          /* [2 trees][2 bits][256] children, then the two roots               */
    struct DecSide *side;          /* per block: what decode_parse_kernel found                        */
    struct DecSideChannel *side_ch;/* per (block, channel)                                             */
    uint32_t num_blocks, pad;
};

/* side information of one channel of one compressed block, as the bitstream carries it */
struct DecSideChannel {
    int32_t head, pre_coef;
    uint32_t order, rshift;
    uint32_t ltp_order, ltp_period;
    int32_t ltp_coef[4];
    int16_t coef[256];             /* coef[i] multiplies x[m - order + i] */
};
struct DecSide { uint32_t status, method; };

/* ---- big-endian bit reader over global memory: 64-bit window, aligned 32-bit refills, one word prefetched.
 * After every operation more than 32 bits are valid, so any field of up to 32 bits is taken from the window without
 * touching memory, and a code whose zero run, stop bit and remainder fit 32 bits is taken from its upper half with
 * 32-bit operations.  The lanes of a warp read different blocks in lockstep, so the refill has NO branch: the word is
 * appended under a predicate and the next one is fetched by a predicated load. ---- */
__device__ __forceinline__ uint32_t dec_load_if(const uint32_t *p, bool take, uint32_t otherwise)
{
    uint32_t v = otherwise;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((uint32_t)take));
    return v;
}

__device__ __forceinline__ void dec_prefetch_if(const uint32_t *p, bool take)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q prefetch.global.L1 [%0];\n\t}" :: "l"(p), "r"((uint32_t)take));
}

struct DecBits {
    const uint32_t *base;          /* aligned word the block's first byte lies in                      */
    uint32_t idx;                  /* next word to fetch, counted from base                            */
    uint32_t limit;                /* first word past the block, counted from base                     */
    uint32_t phase;                /* base's word position inside its 128-byte line                    */
    uint32_t ahead;                /* prefetched word, still in memory byte order                      */
    unsigned long long win;        /* valid bits at the top                                            */
    int avail;
    int skip;                      /* bits of the first word in front of the block                     */
    /* under `take`: the next word (zero behind the block's end) becomes `ahead` */
    __device__ __forceinline__ void fetch_if(bool take)
    {
        const bool inside = take && idx < limit;
        const uint32_t *at = base + idx;
        /* entering a 128-byte line: ask for the one behind it (one lane's miss would stall all the lanes walking with it);
         * a prefetch has no result register, so nothing ever waits for it */
        dec_prefetch_if(at + 32, inside && ((idx + phase) & 31u) == 0u && idx + 32u < limit);
        ahead = dec_load_if(at, inside, take ? 0u : ahead);
        idx += take ? 1u : 0u;
    }
    __device__ __forceinline__ void open(const uint8_t *p, const uint8_t *end)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        base = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        limit = (uint32_t)((((reinterpret_cast<uintptr_t>(end) + 3) & ~(uintptr_t)3) - (a & ~(uintptr_t)3)) >> 2);
        phase = (uint32_t)(a >> 2) & 31u;
        idx = 0u; ahead = 0u;
        skip = (int)(a & 3) * 8;
        dec_prefetch_if(base + (32u - phase), 32u - phase < limit);
        fetch_if(true);
        win = (unsigned long long)__byte_perm(ahead, 0u, 0x0123) << 32;
        win <<= skip;
        avail = 32 - skip;
        fetch_if(true);
        refill();
    }
    /* the reader ran past the end of the block (two words are always in flight): corrupt data */
    __device__ __forceinline__ bool overrun() const { return idx > limit + 3u; }
    /* bits taken from the block so far: everything fetched minus what still waits in the window and the prefetched word */
    __device__ __forceinline__ long long consumed_bits() const { return (long long)idx * 32 - 32 - avail - skip; }
    __device__ __forceinline__ void refill()
    {
        const bool m = avail <= 32;
        const unsigned long long add = (unsigned long long)__byte_perm(ahead, 0u, 0x0123) << ((32 - avail) & 63);
        win |= m ? add : 0ull;
        avail += m ? 32 : 0;
        fetch_if(m);
    }
    __device__ __forceinline__ uint32_t get(uint32_t n)               /* n <= 32; 0 gives 0 (bit_stream.h: GetBits) */
    {
        if (n == 0u) { return 0u; }
        const uint32_t v = (uint32_t)(win >> (64 - n));
        win <<= n; avail -= (int)n;
        refill();
        return v;
    }
    __device__ __forceinline__ uint32_t zero_run()                    /* zeros up to and including the closing 1 */
    {
        uint32_t run = 0;
        for (;;) {
            const int z = __clzll((long long)win);                    /* 64 for an empty window */
            if (z < avail) { win = (win << z) << 1; avail -= z + 1; refill(); return run + (uint32_t)z; }
            run += (uint32_t)avail;
            win = 0; avail = 0;
            refill();
            if (overrun()) { return run; }
        }
    }
};

__device__ __forceinline__ int32_t dec_zigzag(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }

constexpr int kDecStage = 128;               /* samples a warp stages in shared memory for the de-emphasis chain */
constexpr int kDecMaxTaps = 8;                 /* accumulators per lane: 8 x 32 >= order 255 */

struct DecChannel {
    int32_t head, pre_coef;
    uint32_t order, rshift;
    uint32_t ltp_order, ltp_period;
    int32_t ltp_coef[4];
    int32_t cp[32 * kDecMaxTaps + 4];          /* cp[d] multiplies x[m - d], d = 1 .. order; 0 elsewhere */
    int32_t rot[kDecMaxTaps][64];              /* rot[t][i] = cp[(i & 31) + 1 + 32 t]: in step s of a round lane L reads rot[t][31 + L - s],
                                                  a fixed address per lane plus a literal in the unrolled round */
};

/* LPC synthesis of one channel by one warp, T accumulators per lane (order <= 32 T).  Lane L owns the outputs
 * m = L (mod 32); in step s of a round of 32 samples lane s's candidate is final and is broadcast, and every lane
 * adds the product with the tap that lies between that sample and its own next output.  Rounds that touch the
 * warm-up samples (srla_lpc_synthesize.c:253-262) or the end of the block run the general loop; the full rounds
 * behind the warm-up run unrolled: the tap of step s sits at a literal offset from a per-lane address, the owner
 * test compares with a literal.
 * DE: the de-emphasis (srla_utility.c:361-378: y[i] = x[i] + ((y[i-1] c) >> 4), y[-1] = head) rides along -- every lane
 * advances the chain with the broadcast sample (three instructions) and the owner keeps y instead of x; only without a
 * long-term predictor, whose synthesis sits between the two in the reference's order. */
template <int T, bool DE>
__device__ __forceinline__ void dec_lpc_synthesize(int32_t *x, uint32_t n, const DecChannel &chn, uint32_t lane)
{
    const uint32_t order = chn.order, rshift = chn.rshift;
    const uint32_t pc = (uint32_t)chn.pre_coef;
    uint32_t y = (uint32_t)chn.head;
    const uint32_t half = (rshift > 0u) ? (1u << (rshift - 1u)) : 0x80000000u;     /* 1 << -1 on x86 */
    uint32_t acc[T];
    #pragma unroll
    for (int t = 0; t < T; ++t) { acc[t] = half; }                               /* the rounding term rides in the sums */
    uint32_t xprev = 0u, mine = 0u;
    uint32_t next_res = (lane < n) ? (uint32_t)x[lane] : 0u;
    const int32_t *taps = &chn.rot[0][31u + lane];
    for (uint32_t base = 0; base < n; base += 32u) {
        const uint32_t own = base + lane;
        const uint32_t res = next_res;                                              /* this lane's output of the round */
        next_res = (own + 32u < n) ? (uint32_t)x[own + 32u] : 0u;                   /* fetched a round ahead */
        if (base >= order && n - base >= 32u) {
            #pragma unroll
            for (int s = 0; s < 32; ++s) {
                const uint32_t cand = res - (uint32_t)((int32_t)acc[0] >> rshift);
                const uint32_t xq = __shfl_sync(0xffffffffu, cand, s);
                if (DE) { y = xq + (uint32_t)((int32_t)(y * pc) >> 4); }
                if (lane == (uint32_t)s) {
                    mine = DE ? y : xq;
                    #pragma unroll
                    for (int t = 0; t + 1 < T; ++t) { acc[t] = acc[t + 1]; }
                    acc[T - 1] = half;
                }
                #pragma unroll
                for (int t = 0; t < T; ++t) { acc[t] += (uint32_t)taps[64 * t - s] * xq; }
                xprev = xq;
            }
            x[own] = (int32_t)mine;
            continue;
        }
        const uint32_t steps = (n - base < 32u) ? n - base : 32u;
        for (uint32_t s = 0; s < steps; ++s) {
            const uint32_t q = base + s;
            /* every lane evaluates "its" candidate; only the owner's is taken */
            uint32_t cand;
            if (q == 0u) { cand = res; }
            else if (q < order) { cand = res + xprev; }
            else { cand = res - (uint32_t)((int32_t)acc[0] >> rshift); }
            const uint32_t xq = __shfl_sync(0xffffffffu, cand, (int)s);
            if (DE) { y = xq + (uint32_t)((int32_t)(y * pc) >> 4); }
            if (lane == s) {
                mine = DE ? y : xq;
                #pragma unroll
                for (int t = 0; t + 1 < T; ++t) { acc[t] = acc[t + 1]; }
                acc[T - 1] = half;
            }
            xprev = xq;
            const uint32_t d0 = ((lane - s - 1u) & 31u) + 1u;                       /* distance to this lane's next output */
            #pragma unroll
            for (int t = 0; t < T; ++t) { acc[t] += (uint32_t)chn.cp[d0 + 32u * (uint32_t)t] * xq; }
        }
        if (own < n) { x[own] = (int32_t)mine; }
    }
    __syncwarp();
}

/* The serial walk over a compressed block (srla_decoder.c:436-540, srla_coder.c:596-690).  A block's bitstream is one
 * chain of dependent operations per code (window -> zero run -> code length -> shifted window -> refill), and every
 * block is walked by ONE thread; the lanes of a warp walk different blocks in LOCKSTEP: one flat loop whose trip is one
 * code, taken from the upper half of the window without a branch (zero run, stop bit and remainder in at most 32
 * bits; anything longer takes the general path).  Partition parameters and channel heads are read in a rarely taken
 * branch -- partitions are power-of-two fractions of a block, so the lanes mostly take it in the same trip.  Side
 * information goes to p.side / p.side_ch, residuals straight into the output buffer.  Header problems (sync, size,
 * checksum, type) are judged by decode_blocks_kernel; blocks that are not well-formed compressed blocks are skipped
 * here.
 * Measured and dropped (round 2, config-2 stream, parse kernel alone): two blocks per thread with a branch-free trip for
 * both (the two chains interleave in the SASS, but a trip then costs twice the instructions and takes 720 instead of
 * 415 clocks: 3.0 ms against 1.7); a walk that keeps only a bit position and loads the two words under it for every
 * code (40 % fewer instructions per code, but an L1 load on every trip of the chain and a reader to open between the
 * partitions: 2.2 ms). */
struct DecWalk {
    DecBits br;
    int32_t *out, *x;              /* the block's first channel / the channel being filled             */
    uint32_t n, payload_bits;
    uint32_t ch, i, in_part, parts, k, per, rec;
    uint32_t status, method;
    bool first, open, done, walked;
};

/* sync code, sizes, side information of every channel (srla_decoder.c:436-540); leaves the reader in front of the
 * first channel's residual codes */
__device__ __forceinline__ void dec_walk_head(DecWalk &w, const DecParams &p, uint32_t bi)
{
    w.status = 0u; w.method = 0u; w.done = true; w.walked = false;
    w.ch = 0u; w.i = 0u; w.in_part = 0u; w.parts = 0u; w.k = 0u; w.per = 0u; w.rec = 0u; w.first = false; w.open = false;
    w.n = 0u; w.payload_bits = 0u; w.out = p.out; w.x = p.out;
    if (bi >= p.num_blocks) { return; }
    const DecBlock blk = p.blocks[bi];
    const uint8_t *b = p.data + blk.offset;
    const uint32_t nch = p.nch, n = blk.nsmpl;
    const uint32_t size = ((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5];
    const bool walk = b[0] == 0xFFu && b[1] == 0xFFu && size >= 5u && size + 6u <= blk.bytes && b[8] == (uint8_t)kBlockCompress
                      && ((((uint32_t)b[9] << 8) | b[10]) == n) && n > 0u;
    if (!walk) { return; }
    const uint8_t *payload = b + 11;
    const uint32_t payload_bytes = blk.bytes - 11u;
    w.walked = true; w.n = n; w.payload_bits = payload_bytes * 8u;
    w.out = p.out + blk.sample_offset; w.x = w.out;
    DecSideChannel *chan = p.side_ch + (size_t)bi * nch;
    DecBits &br = w.br;
    br.open(payload, payload + payload_bytes);
    const uint16_t *tree0 = p.tree, *tree1 = p.tree + 512;
    const uint32_t root0 = p.tree[1024], root1 = p.tree[1025];
    uint32_t status = 0u;
    w.method = br.get(2);
    for (uint32_t ch = 0; ch < nch; ++ch) {
        chan[ch].head = dec_zigzag(br.get(p.bps + 1u));
        chan[ch].pre_coef = dec_zigzag(br.get(5));
    }
    for (uint32_t ch = 0; ch < nch && !status; ++ch) {
        DecSideChannel &c = chan[ch];
        const uint32_t order = br.get(8);
        c.order = order; c.rshift = br.get(4);
        const uint32_t use_sum = br.get(1);
        int32_t prev = 0;
        for (uint32_t i = 0; i < order; ++i) {
            const uint16_t *tree = (use_sum && i > 0u) ? tree1 : tree0;
            uint32_t node = (use_sum && i > 0u) ? root1 : root0;
            do { node = tree[br.get(1) * 256u + (node - 256u)]; } while (node >= 256u);
            int32_t v = dec_zigzag(node & 255u);
            if (use_sum && i > 0u) { v -= prev; }                                /* summed-neighbour table: coef[i] = code - coef[i-1] */
            prev = v;
            c.coef[i] = (int16_t)v;
        }
        if (br.overrun()) { status = SRLA_APIRESULT_DETECT_DATA_CORRUPTION; }
    }
    for (uint32_t ch = 0; ch < nch; ++ch) {
        DecSideChannel &c = chan[ch];
        uint32_t ltp_order = 0u, ltp_period = 0u;
        if (br.get(1)) {
            ltp_order = 2u * br.get(1) + 1u;
            ltp_period = br.get(8) + (uint32_t)kLtpMinPeriod;
            for (uint32_t i = 0; i < ltp_order; ++i) { c.ltp_coef[i] = dec_zigzag(br.get(6)); }
        }
        c.ltp_order = ltp_order; c.ltp_period = ltp_period;
    }
    w.status = status;
    w.done = status != 0u;
}

/* Between two partitions (srla_coder.c:648-690): the next partition's parameter, in front of it the next channel's
 * code type and partition order, behind a channel's last partition the samples no partition covers (never for streams
 * the encoder writes).  Returns with codes to read (in_part > 0) or with the walk finished. */
__device__ __forceinline__ void dec_walk_between(DecWalk &w, const DecParams &p)
{
    DecBits &br = w.br;
    const uint32_t n = w.n;
    for (;;) {
        if (w.parts == 0u) {
            if (w.open) { for (; w.i < n; ++w.i) { w.x[w.i] = 0; } ++w.ch; w.open = false; }
            if (w.ch >= p.nch) { w.done = true; return; }
            w.x = w.out + (size_t)w.ch * p.stride; w.i = 0u;
            const uint32_t code = br.get(2);
            if (code == (uint32_t)kCodeAllZero) { for (uint32_t j = 0; j < n; ++j) { w.x[j] = 0; } ++w.ch; continue; }
            if (code > (uint32_t)kCodeAllZero) { w.status = SRLA_APIRESULT_DETECT_DATA_CORRUPTION; w.done = true; return; }
            const uint32_t porder = br.get(10);
            if (porder > (uint32_t)kLog2MaxParts) { w.status = SRLA_APIRESULT_DETECT_DATA_CORRUPTION; w.done = true; return; }
            w.per = n >> porder; w.parts = 1u << porder; w.first = true; w.open = true;
            w.rec = (code == (uint32_t)kCodeRecursiveRice) ? 1u : 0u;
        }
        if (br.overrun()) { w.status = SRLA_APIRESULT_DETECT_DATA_CORRUPTION; w.done = true; return; }
        w.k = w.first ? br.get(5) : (uint32_t)((int32_t)w.k + dec_zigzag(br.zero_run()));
        w.first = false; --w.parts;
        if (w.k > 31u) { w.status = SRLA_APIRESULT_DETECT_DATA_CORRUPTION; w.done = true; return; }
        w.in_part = w.per;
        if (w.per != 0u) { return; }                                                 /* else more partitions than samples: parameters only */
    }
}

/* One code: Rice = unary quotient + k bits; recursive Rice = k + 1 bits behind an empty run, else k bits and the
 * quotient one up (srla_coder.c:610-646).  dec_code_shape looks at the upper half of the window (all valid); a code
 * of at most 32 bits is then taken without a branch. */
struct DecShape { uint32_t hi, z, nb, need; };
__device__ __forceinline__ DecShape dec_code_shape(const DecWalk &w)
{
    DecShape s;
    s.hi = (uint32_t)(w.br.win >> 32);
    s.z = (uint32_t)__clz((int)s.hi);                                                /* 32 for an empty half */
    s.nb = w.k + (w.rec & (s.z == 0u ? 1u : 0u));
    s.need = s.z + 1u + s.nb;
    return s;
}
__device__ __forceinline__ void dec_put(DecWalk &w, uint32_t quot, uint32_t low)
{
    const uint32_t u = low + ((quot + (w.rec & (quot ? 1u : 0u))) << w.k);           /* low < 2^k whenever the quotient counts */
    w.x[w.i] = dec_zigzag(u);
    ++w.i; --w.in_part;
}
__device__ __forceinline__ void dec_code_short(DecWalk &w, const DecShape &s)         /* s.need <= 32 */
{
    const uint32_t low = __funnelshift_rc(__funnelshift_lc(0u, s.hi, s.z + 1u), 0u, 32u - s.nb);   /* (hi << (z + 1)) >> (32 - nb), shifts of 32 give 0 */
    w.br.win = (w.br.win << (s.need - 1u)) << 1; w.br.avail -= (int)s.need;
    w.br.refill();
    dec_put(w, s.z, low);
}
__device__ __forceinline__ void dec_code_any(DecWalk &w)
{
    const DecShape s = dec_code_shape(w);
    if (s.need <= 32u) { dec_code_short(w, s); return; }
    const uint32_t quot = w.br.zero_run();
    const uint32_t low = w.br.get(w.k + (w.rec & (quot == 0u ? 1u : 0u)));
    dec_put(w, quot, low);
}

__global__ void __launch_bounds__(32) decode_parse_kernel(const DecParams p)
{
    const uint32_t bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= p.num_blocks) { return; }
    DecWalk w;
    dec_walk_head(w, p, bi);
    while (!w.done) {
        if (w.in_part == 0u) { dec_walk_between(w, p); continue; }
        dec_code_any(w);
    }
    /* exact bound: a walk that took more bits than the block holds read its neighbour's bytes (corrupt data) */
    if (w.walked && !w.status && w.br.consumed_bits() > (long long)w.payload_bits) { w.status = SRLA_APIRESULT_DETECT_DATA_CORRUPTION; }
    DecSide sd; sd.status = w.status; sd.method = w.method;
    p.side[bi] = sd;
}

__global__ void __launch_bounds__(256) decode_blocks_kernel(const DecParams p)
{
    extern __shared__ __align__(16) unsigned char dec_smem[];
    DecChannel *chan = reinterpret_cast<DecChannel *>(dec_smem);                       /* [channels] */
    int32_t *stage = reinterpret_cast<int32_t *>(chan + p.nch);                        /* [channels][kDecStage] */
    __shared__ uint32_t sh_type, sh_n, sh_status, sh_method;
    __shared__ uint32_t sh_sum[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const DecBlock blk = p.blocks[blockIdx.x];
    const uint8_t *b = p.data + blk.offset;
    const uint32_t nch = p.nch;
    if (tid == 0u) { sh_status = 0u; sh_type = 3u; sh_n = 0u; sh_method = 0u; }
    __syncthreads();

    /* ---- block header (srla_decoder.c:661-700): sync, size, checksum, type, sample count ---- */
    if (warp == 0u) {
        const uint32_t size = ((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5];
        const bool size_ok = size >= 5u && size + 6u <= blk.bytes;
        if (lane == 0u) { sh_sum[0] = 0u; sh_sum[1] = 0u; }
        __syncwarp();
        if (p.check && size_ok) {
            /* Fletcher-16 over the size - 2 bytes after the checksum field: lanes take contiguous chunks,
             * sum2 of the whole = sum over chunks of (sum2_chunk + bytes_after_chunk_start ... ) folded below */
            const uint32_t len = size - 2u;
            const uint8_t *q = b + 8;
            const uint32_t per = (len + 31u) / 32u;
            const uint32_t lo = min(len, lane * per), hi = min(len, lo + per);
            uint32_t s1 = 0, s2 = 0;
            for (uint32_t i = lo; i < hi; ++i) {
                s1 += q[i]; s2 += s1;
                if (((i - lo) & 4095u) == 4095u) { s1 %= 255u; s2 %= 255u; }
            }
            s1 %= 255u; s2 %= 255u;
            /* combine in lane order: after a chunk of length L, sum2 += L * sum1_before */
            uint32_t t1 = 0, t2 = 0;
            for (int l = 0; l < 32; ++l) {
                const uint32_t c1 = __shfl_sync(0xffffffffu, s1, l), c2 = __shfl_sync(0xffffffffu, s2, l);
                const uint32_t cl = min(len, min(len, (uint32_t)l * per) + per) - min(len, (uint32_t)l * per);
                t2 = (t2 + c2 + (cl % 255u) * t1) % 255u;
                t1 = (t1 + c1) % 255u;
            }
            if (lane == 0u) { sh_sum[0] = t1; sh_sum[1] = t2; }
        }
        __syncwarp();
        if (lane == 0u) {
            uint32_t status = 0u;
            if (b[0] != 0xFFu || b[1] != 0xFFu) { status = SRLA_APIRESULT_INVALID_FORMAT; }
            else if (size + 6u > blk.bytes) { status = SRLA_APIRESULT_INSUFFICIENT_DATA; }
            else if (size < 5u) { status = SRLA_APIRESULT_INVALID_FORMAT; }
            else if (p.check && ((sh_sum[1] << 8) | sh_sum[0]) != (((uint32_t)b[6] << 8) | b[7])) { status = SRLA_APIRESULT_DETECT_DATA_CORRUPTION; }
            else if (b[8] > 2u) { status = SRLA_APIRESULT_INVALID_FORMAT; }
            sh_type = b[8];
            sh_n = ((uint32_t)b[9] << 8) | b[10];
            if (!status && sh_n != blk.nsmpl) { status = SRLA_APIRESULT_INVALID_FORMAT; }
            sh_status = status;
        }
    }
    __syncthreads();
    const uint32_t n = sh_n, type = sh_type;
    int32_t *out = p.out + blk.sample_offset;
    if (sh_status != 0u) { if (tid == 0u) { p.status[blockIdx.x] = sh_status; } return; }
    const uint8_t *payload = b + 11;
    const uint32_t payload_bytes = blk.bytes - 11u;

    if (type == (uint32_t)kBlockSilent) {
        for (uint32_t i = tid; i < n * nch; i += blockDim.x) { out[(size_t)(i / n) * p.stride + (i % n)] = 0; }
        if (tid == 0u) { p.status[blockIdx.x] = 0u; }
        return;
    }
    if (type == (uint32_t)kBlockRaw) {
        /* srla_decoder.c:363-433: frames of big-endian zig-zag samples */
        const uint32_t sb = p.bps >> 3;
        if ((uint64_t)payload_bytes < ((uint64_t)p.bps * n * nch) / 8u) { if (tid == 0u) { p.status[blockIdx.x] = SRLA_APIRESULT_INSUFFICIENT_DATA; } return; }
        for (uint32_t i = tid; i < n * nch; i += blockDim.x) {
            const uint8_t *q = payload + (size_t)i * sb;
            uint32_t u = 0;
            for (uint32_t k = 0; k < sb; ++k) { u = (u << 8) | q[k]; }
            const uint32_t smpl = i / nch, ch = i - smpl * nch;
            out[(size_t)ch * p.stride + smpl] = dec_zigzag(u);
        }
        if (tid == 0u) { p.status[blockIdx.x] = 0u; }
        return;
    }

    /* ---- compressed block: decode_parse_kernel has walked the bitstream ---- */
    if (n == 0u) {
        /* a block that announces no samples (only a hostile or broken stream has one): decode_parse_kernel skips it and
         * leaves its side records unwritten, so they must not be read; there is nothing to synthesise */
        if (tid == 0u) { p.status[blockIdx.x] = SRLA_APIRESULT_INVALID_FORMAT; }
        return;
    }
    {
        const DecSide sd = p.side[blockIdx.x];
        if (sd.status != 0u) { if (tid == 0u) { p.status[blockIdx.x] = sd.status; } return; }
        if (tid == 0u) { sh_method = sd.method; }
        if (warp < nch) {
            const DecSideChannel &g = p.side_ch[(size_t)blockIdx.x * nch + warp];
            DecChannel &c = chan[warp];
            for (uint32_t d = lane; d < 32u * kDecMaxTaps + 4u; d += 32u) { c.cp[d] = 0; }
            __syncwarp();
            const uint32_t order = min(g.order, (uint32_t)kMaxOrder);      /* the order field is 8 bits; never index past cp[] */
            for (uint32_t i = lane; i < order; i += 32u) { c.cp[order - i] = g.coef[i]; }
            __syncwarp();
            for (uint32_t i = lane; i < 64u * kDecMaxTaps; i += 32u) { c.rot[i >> 6][i & 63u] = c.cp[(i & 31u) + 1u + 32u * (i >> 6)]; }
            if (lane == 0u) {
                c.head = g.head; c.pre_coef = g.pre_coef; c.order = order; c.rshift = g.rshift;
                c.ltp_order = g.ltp_order; c.ltp_period = g.ltp_period;
                for (int i = 0; i < 4; ++i) { c.ltp_coef[i] = g.ltp_coef[i]; }
            }
        }
    }
    __syncthreads();

    /* ---- synthesis: warp w = channel w ---- */
    if (warp < nch && n > 0u) {
        int32_t *x = out + (size_t)warp * p.stride;
        const DecChannel &c = chan[warp];
        const bool ltp = c.ltp_period > 0u && c.ltp_order > 0u;
        bool deemphasised = false;
        if (c.order > 0u && n > c.order) {
            if (!ltp && c.order <= 32u) { dec_lpc_synthesize<1, true>(x, n, c, lane); deemphasised = true; }
            else if (!ltp && c.order <= 64u) { dec_lpc_synthesize<2, true>(x, n, c, lane); deemphasised = true; }
            else if (c.order <= 32u) { dec_lpc_synthesize<1, false>(x, n, c, lane); }
            else if (c.order <= 64u) { dec_lpc_synthesize<2, false>(x, n, c, lane); }
            else if (c.order <= 128u) { dec_lpc_synthesize<4, false>(x, n, c, lane); }
            else { dec_lpc_synthesize<8, false>(x, n, c, lane); }
        } else if (c.order > 0u) {
            /* a block no longer than the order is all warm-up (srla_lpc_synthesize.c:253-255): running sum */
            if (lane == 0u) { for (uint32_t i = 1; i < n && i < c.order; ++i) { x[i] = (int32_t)((uint32_t)x[i] + (uint32_t)x[i - 1]); } }
            __syncwarp();
        }
        if (ltp) {
            /* srla_lpc_synthesize.c:264-327: x[s] += (16 + sum_j c[j] x[s - T - h + j]) >> 5 for s > T + h; a chunk shorter
             * than the shortest lag T - h never reads what it writes */
            const uint32_t h = c.ltp_order >> 1, T0 = c.ltp_period;
            const uint32_t chunk = min(32u, T0 - h);
            for (uint32_t s0 = T0 + h + 1u; s0 < n; s0 += chunk) {
                const uint32_t s = s0 + lane;
                if (lane < chunk && s < n) {
                    uint32_t predict = 16u;
                    for (uint32_t j = 0; j < c.ltp_order; ++j) { predict += (uint32_t)c.ltp_coef[j] * (uint32_t)x[s - T0 - h + j]; }
                    x[s] = (int32_t)((uint32_t)x[s] + (uint32_t)((int32_t)predict >> 5));
                }
                __syncwarp();
            }
        }
        if (!deemphasised) {
            /* de-emphasis (srla_utility.c:361-378): x[0] += (head c) >> 4, x[i] += (x[i-1] c) >> 4 -- a serial chain.  The
             * warp moves 128 samples at a time through shared memory (coalesced both ways); lane 0 runs the chain there.
             * (Blocks without a long-term predictor and an order of at most 64 had it done inside the synthesis rounds.) */
            const uint32_t pc = (uint32_t)c.pre_coef;
            int32_t *st = stage + warp * kDecStage;
            uint32_t prev = (uint32_t)c.head;
            for (uint32_t i0 = 0; i0 < n; i0 += (uint32_t)kDecStage) {
                const uint32_t len = min((uint32_t)kDecStage, n - i0);
                for (uint32_t k = lane; k < len; k += 32u) { st[k] = x[i0 + k]; }
                __syncwarp();
                if (lane == 0u) {
                    #pragma unroll 8
                    for (uint32_t k = 0; k < len; ++k) { prev = (uint32_t)st[k] + (uint32_t)((int32_t)(prev * pc) >> 4); st[k] = (int32_t)prev; }
                }
                __syncwarp();
                for (uint32_t k = lane; k < len; k += 32u) { x[i0 + k] = st[k]; }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    /* ---- channel reconstruction (srla_utility.c:106-174) and the offset shift (srla_decoder.c:582-590) ---- */
    const uint32_t method = sh_method, sh = p.lshift;
    if (nch >= 2u && method != 0u) {
        int32_t *c0 = out, *c1 = out + p.stride;
        for (uint32_t i = tid; i < n; i += blockDim.x) {
            uint32_t a = (uint32_t)c0[i], s = (uint32_t)c1[i];
            if (method == 1u) { a -= (uint32_t)((int32_t)s >> 1); s += a; }          /* MS */
            else if (method == 2u) { s += a; }                                       /* LS */
            else { a = s - a; }                                                      /* SR */
            c0[i] = (int32_t)(a << sh); c1[i] = (int32_t)(s << sh);
        }
        if (sh > 0u) { for (uint32_t i = tid; i < n * (nch - 2u); i += blockDim.x) { int32_t *v = out + (size_t)(2u + i / n) * p.stride + (i % n); *v = (int32_t)((uint32_t)*v << sh); } }
    } else if (sh > 0u) {
        for (uint32_t i = tid; i < n * nch; i += blockDim.x) { int32_t *v = out + (size_t)(i / n) * p.stride + (i % n); *v = (int32_t)((uint32_t)*v << sh); }
    }
    if (tid == 0u) { p.status[blockIdx.x] = 0u; }
}

} // namespace srla

/* ================================================================================================
 * host side: the reference's decoder API
 * ============================================================================================== */
namespace {

const uint32_t kDecoderMagic = 0x53424C44u;

struct DecoderCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaStream_t lanes[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };   /* a block's walk is ~2.5 ms of serial latency whatever the
                                                       launch size: the groups of a pipelined call must run side by side */
    cudaEvent_t ev_table = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k0, ev_k1, ev_out;      /* per group of a pipelined call */
    PinBuf h_in, h_out;
    std::unique_ptr<WorkerPool> pool;
    int host_threads = 8;
    int pipeline = -1;             /* SRLA_B200_DECODE_PIPELINE: 1 always, 0 never, default: from a handle's second long stream on */
    int long_calls = 0;
    DevBuf data, out, blocks, status, tree, side, side_ch;
    int parse_lanes = 0;           /* SRLA_B200_DECODE_LANES: blocks walked per warp of decode_parse_kernel; 0 = from the launch size */
    int sms = 148;
    PinBuf h_blocks, h_status;
    float last_ms = 0.f;
};

} // namespace

struct SRLADecoder {
    uint32_t magic;
    struct SRLADecoderConfig config;
    struct SRLAHeader header;
    int set_header;
    uint8_t alloced_by_own;
    void *work;
    DecoderCtx *ctx;
};

namespace {

/* Blocks per warp of decode_parse_kernel.  The walk is a chain of dependent operations per code, so a launch wants
 * a few warps on every scheduler before it fills the lanes of a warp (lanes in lockstep wait for each other's rare
 * slow paths): about kParseWarpsPerSm warps per SM, then up to 32 lanes. */
constexpr size_t kParseWarpsPerSm = 8;
unsigned parse_lanes_for(const DecoderCtx *c, size_t blocks)
{
    if (c->parse_lanes > 0) { return (unsigned)c->parse_lanes; }
    const size_t warps = (size_t)c->sms * kParseWarpsPerSm;
    return (unsigned)std::min<size_t>(32, std::max<size_t>(1, (blocks + warps - 1) / warps));
}

bool decoder_ctx_init(DecoderCtx *c)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        std::fprintf(stderr, "[srla_b200] no CUDA device: the SRLA B200 decode path has no CPU fallback\n");
        return false;
    }
    if (g_device >= 0) { CU_TRY(cudaSetDevice(g_device)); }
    CU_TRY(cudaGetDevice(&c->device));
    if (!device_is_sm100(c->device)) { return false; }
    CU_TRY(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, c->device));
    CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreate(&c->ev0));
    CU_TRY(cudaEventCreate(&c->ev1));
    CU_TRY(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    for (cudaStream_t &l : c->lanes) { CU_TRY(cudaStreamCreateWithFlags(&l, cudaStreamNonBlocking)); }
    CU_TRY(cudaEventCreateWithFlags(&c->ev_table, cudaEventDisableTiming));
    c->host_threads = (int)std::max(2u, std::min(16u, usable_cpus()));
    if (const char *e = std::getenv("SRLA_B200_FEED_THREADS")) { const int v = std::atoi(e); if (v >= 0 && v <= 64) { c->host_threads = v; } }
    if (const char *e = std::getenv("SRLA_B200_DECODE_PIPELINE")) { c->pipeline = std::atoi(e) ? 1 : 0; }
    if (const char *e = std::getenv("SRLA_B200_DECODE_LANES")) { const int v = std::atoi(e); if (v >= 0 && v <= 32) { c->parse_lanes = v; } }
    host::HuffTable plain, summed; host::HuffTree t0, t1;
    host::build_format_huffman(plain, summed, &t0, &t1);
    uint16_t tab[1026];
    for (int b = 0; b < 2; b++) { for (int i = 0; i < 256; i++) { tab[b * 256 + i] = t0.child[b][i]; tab[512 + b * 256 + i] = t1.child[b][i]; } }
    tab[1024] = t0.root; tab[1025] = t1.root;
    if (!c->tree.reserve(sizeof(tab))) { return false; }
    CU_TRY(cudaMemcpy(c->tree.p, tab, sizeof(tab), cudaMemcpyHostToDevice));
    return true;
}

void decoder_ctx_destroy(DecoderCtx *c)
{
    if (!c) { return; }
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->ev0) { cudaEventDestroy(c->ev0); }
    if (c->ev1) { cudaEventDestroy(c->ev1); }
    c->pool.reset();
    if (c->copy_in) { cudaStreamSynchronize(c->copy_in); cudaStreamDestroy(c->copy_in); }
    if (c->copy_out) { cudaStreamSynchronize(c->copy_out); cudaStreamDestroy(c->copy_out); }
    for (cudaStream_t &l : c->lanes) { if (l) { cudaStreamSynchronize(l); cudaStreamDestroy(l); } }
    if (c->ev_table) { cudaEventDestroy(c->ev_table); }
    for (std::vector<cudaEvent_t> *v : { &c->ev_in, &c->ev_k0, &c->ev_k1, &c->ev_out }) { for (cudaEvent_t e : *v) { cudaEventDestroy(e); } }
    c->h_in.release(); c->h_out.release();
    DevBuf *bufs[] = { &c->data, &c->out, &c->blocks, &c->status, &c->tree, &c->side, &c->side_ch };
    for (DevBuf *b : bufs) { b->release(); }
    c->h_blocks.release(); c->h_status.release();
}

SRLAApiResult decoder_header_valid(const struct SRLAHeader *h)            /* srla_decoder.c:137-182 */
{
    if (h->format_version != SRLA_FORMAT_VERSION || h->codec_version != SRLA_CODEC_VERSION) { return SRLA_APIRESULT_INVALID_FORMAT; }
    if (h->num_channels == 0 || h->num_samples == 0 || h->sampling_rate == 0 || h->bits_per_sample == 0) { return SRLA_APIRESULT_INVALID_FORMAT; }
    if (h->offset_lshift >= 32 || h->max_num_samples_per_block == 0 || h->preset >= SRLA_NUM_PARAMETER_PRESETS) { return SRLA_APIRESULT_INVALID_FORMAT; }
    return SRLA_APIRESULT_OK;
}

/* Long streams: the blocks are cut into groups whose stages overlap -- host threads stage the group's bytes in
 * page-locked memory, the copy engine moves them in, the two kernels decode the group, the copy engine moves its
 * samples out to page-locked memory and the host threads pass them on to the caller's (pageable) channel buffers.
 * A group is only passed on once every block in it has decoded cleanly, so that -- like the reference, which stops
 * at the first bad block (srla_decoder.c:780-786) -- nothing behind a bad block is delivered. */
SRLAApiResult decoder_run_pipelined(struct SRLADecoder *d, const uint8_t *data, uint64_t data_bytes, const std::vector<DecBlock> &blocks,
                                    int32_t **buffer, uint32_t total_samples)
{
    DecoderCtx *c = d->ctx;
    const uint32_t nch = d->header.num_channels;
    const uint64_t stride = round_up_u32(total_samples, 16);
    const size_t nb = blocks.size();
    const size_t G = std::min<size_t>(12, std::max<size_t>(2, nb / 128));
    if (!c->data.reserve(data_bytes + 16) || !c->out.reserve(sizeof(int32_t) * stride * nch) || !c->blocks.reserve(sizeof(DecBlock) * nb)
        || !c->status.reserve(sizeof(uint32_t) * nb) || !c->side.reserve(sizeof(DecSide) * nb) || !c->side_ch.reserve(sizeof(DecSideChannel) * nb * nch)
        || !c->h_blocks.reserve(sizeof(DecBlock) * nb) || !c->h_status.reserve(sizeof(uint32_t) * nb)
        || !c->h_in.reserve(data_bytes + 16) || !c->h_out.reserve(sizeof(int32_t) * stride * nch)) { return SRLA_APIRESULT_NG; }
    for (std::vector<cudaEvent_t> *v : { &c->ev_in, &c->ev_out }) {
        while (v->size() < G) { cudaEvent_t e; if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return SRLA_APIRESULT_NG; } v->push_back(e); }
    }
    for (std::vector<cudaEvent_t> *v : { &c->ev_k0, &c->ev_k1 }) {
        while (v->size() < G) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) { return SRLA_APIRESULT_NG; } v->push_back(e); }
    }
    if (!c->pool) { c->pool.reset(new WorkerPool()); }
    c->pool->ensure(c->host_threads);

    /* group boundaries (blocks), their byte and sample ranges */
    std::vector<size_t> gb(G + 1);
    for (size_t g = 0; g <= G; g++) { gb[g] = nb * g / G; }
    auto byte_begin = [&](size_t g) { return (uint64_t)blocks[gb[g]].offset; };
    auto byte_end = [&](size_t g) { return (uint64_t)blocks[gb[g + 1] - 1].offset + blocks[gb[g + 1] - 1].bytes; };
    auto smp_begin = [&](size_t g) { return blocks[gb[g]].sample_offset; };
    auto smp_end = [&](size_t g) { return blocks[gb[g + 1] - 1].sample_offset + blocks[gb[g + 1] - 1].nsmpl; };

    struct Item { void *dst; const void *src; size_t bytes; uint32_t group; };
    std::vector<Item> in_items, out_items;
    std::unique_ptr<std::atomic<int>[]> in_left(new std::atomic<int>[G]);
    constexpr size_t kChunk = 1u << 20;
    for (size_t g = 0; g < G; g++) {
        in_left[g].store(0);
        for (uint64_t at = byte_begin(g); at < byte_end(g); at += kChunk) {
            in_items.push_back({ (uint8_t *)c->h_in.p + at, data + at, (size_t)std::min<uint64_t>(kChunk, byte_end(g) - at), (uint32_t)g });
            in_left[g].fetch_add(1);
        }
        for (uint32_t ch = 0; ch < nch; ch++) {
            for (uint64_t at = smp_begin(g); at < smp_end(g); at += kChunk / 4) {
                const size_t cnt = (size_t)std::min<uint64_t>(kChunk / 4, smp_end(g) - at);
                out_items.push_back({ buffer[ch] + at, (const int32_t *)c->h_out.p + stride * ch + at, cnt * sizeof(int32_t), (uint32_t)g });
            }
        }
    }
    std::atomic<size_t> in_next{0}, out_next{0}, out_ready{0}, out_limit{G};
    c->pool->start([&] {
        for (;;) {
            const size_t i = in_next.fetch_add(1);
            if (i >= in_items.size()) { break; }
            std::memcpy(in_items[i].dst, in_items[i].src, in_items[i].bytes);
            in_left[in_items[i].group].fetch_sub(1, std::memory_order_release);
        }
        for (;;) {
            const size_t i = out_next.fetch_add(1);
            if (i >= out_items.size()) { break; }
            const Item &it = out_items[i];
            for (;;) {
                if (out_ready.load(std::memory_order_acquire) > it.group) { std::memcpy(it.dst, it.src, it.bytes); break; }
                if (out_limit.load(std::memory_order_acquire) <= it.group) { break; }        /* behind a bad block: never delivered */
                std::this_thread::yield();
            }
        }
    });
    struct Finish { WorkerPool *pool; std::atomic<size_t> *limit, *ready; ~Finish() { limit->store(ready->load()); pool->wait(); } } finish{ c->pool.get(), &out_limit, &out_ready };

    std::memcpy(c->h_blocks.p, blocks.data(), sizeof(DecBlock) * nb);
    if (cudaMemcpyAsync(c->blocks.p, c->h_blocks.p, sizeof(DecBlock) * nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess
        || cudaMemsetAsync((uint8_t *)c->data.p + data_bytes, 0, 16, c->stream) != cudaSuccess
        || cudaMemsetAsync(c->status.p, 0xff, sizeof(uint32_t) * nb, c->stream) != cudaSuccess
        || cudaMemsetAsync(c->side.p, 0, sizeof(DecSide) * nb, c->stream) != cudaSuccess
        || cudaMemsetAsync(c->side_ch.p, 0, sizeof(DecSideChannel) * nb * nch, c->stream) != cudaSuccess
        || cudaEventRecord(c->ev_table, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    for (cudaStream_t l : c->lanes) { if (cudaStreamWaitEvent(l, c->ev_table, 0) != cudaSuccess) { return SRLA_APIRESULT_NG; } }
    DecParams base;
    base.data = (const uint8_t *)c->data.p; base.out = (int32_t *)c->out.p; base.stride = stride;
    base.nch = nch; base.bps = d->header.bits_per_sample; base.lshift = d->header.offset_lshift; base.check = (d->config.check_checksum == 1) ? 1u : 0u;
    base.tree = (const uint16_t *)c->tree.p; base.pad = 0;

    /* passes finished groups on, in order; stops for good at the first group with a bad block */
    size_t published = 0, issued = 0; bool failed = false; long bad_block = -1;
    const uint32_t *st = (const uint32_t *)c->h_status.p;
    auto publish = [&](bool block) {
        while (published < issued && bad_block < 0 && !failed) {
            if (block) { if (cudaEventSynchronize(c->ev_out[published]) != cudaSuccess) { failed = true; break; } }
            else {
                const cudaError_t q = cudaEventQuery(c->ev_out[published]);
                if (q == cudaErrorNotReady) { (void)cudaGetLastError(); break; }
                if (q != cudaSuccess) { failed = true; break; }
            }
            for (size_t i = gb[published]; i < gb[published + 1]; i++) { if (st[i] != 0u) { bad_block = (long)i; break; } }
            if (bad_block >= 0) { out_limit.store(published, std::memory_order_release); break; }
            published++;
            out_ready.store(published, std::memory_order_release);
        }
    };
    for (size_t g = 0; g < G; g++) {
        while (in_left[g].load(std::memory_order_acquire) > 0) { publish(false); std::this_thread::yield(); }
        const size_t b0 = gb[g], cnt = gb[g + 1] - gb[g];
        if (cudaMemcpyAsync((uint8_t *)c->data.p + byte_begin(g), (const uint8_t *)c->h_in.p + byte_begin(g), byte_end(g) - byte_begin(g), cudaMemcpyHostToDevice, c->copy_in) != cudaSuccess
            || cudaEventRecord(c->ev_in[g], c->copy_in) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        const cudaStream_t on = c->lanes[g % 6];
        if (cudaStreamWaitEvent(on, c->ev_in[g], 0) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        DecParams p = base;
        p.blocks = (const DecBlock *)c->blocks.p + b0; p.status = (uint32_t *)c->status.p + b0;
        p.side = (DecSide *)c->side.p + b0; p.side_ch = (DecSideChannel *)c->side_ch.p + b0 * nch; p.num_blocks = (uint32_t)cnt;
        cudaEventRecord(c->ev_k0[g], on);
        const unsigned pl = parse_lanes_for(c, nb);                 /* the groups' kernels run side by side: the call's blocks count */
        decode_parse_kernel<<<(unsigned)((cnt + pl - 1) / pl), pl, 0, on>>>(p);
        decode_blocks_kernel<<<(unsigned)cnt, 32u * std::max(1u, nch), nch * (sizeof(DecChannel) + sizeof(int32_t) * kDecStage), on>>>(p);
        cudaEventRecord(c->ev_k1[g], on);
        if (cudaGetLastError() != cudaSuccess || cudaStreamWaitEvent(c->copy_out, c->ev_k1[g], 0) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        if (cudaMemcpyAsync((uint32_t *)c->h_status.p + b0, (const uint32_t *)c->status.p + b0, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, c->copy_out) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        for (uint32_t ch = 0; ch < nch; ch++) {
            const size_t off = stride * ch + smp_begin(g);
            if (cudaMemcpyAsync((int32_t *)c->h_out.p + off, (const int32_t *)c->out.p + off, sizeof(int32_t) * (smp_end(g) - smp_begin(g)), cudaMemcpyDeviceToHost, c->copy_out) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        }
        if (cudaEventRecord(c->ev_out[g], c->copy_out) != cudaSuccess) { return SRLA_APIRESULT_NG; }
        issued = g + 1;
        publish(false);
    }
    publish(true);
    bool lanes_ok = true;
    for (cudaStream_t l : c->lanes) { lanes_ok = lanes_ok && cudaStreamSynchronize(l) == cudaSuccess; }
    if (cudaStreamSynchronize(c->copy_out) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess || !lanes_ok || failed) {
        std::fprintf(stderr, "[srla_b200] decode failed on the device: %s\n", cudaGetErrorString(cudaGetLastError()));
        return SRLA_APIRESULT_NG;
    }
    /* device time from the first group's kernels to the last group's (they overlap each other and the copies) */
    c->last_ms = 0.f;
    for (size_t g = 0; g < G; g++) { float t = 0.f; cudaEventElapsedTime(&t, c->ev_k0[0], c->ev_k1[g]); c->last_ms = std::max(c->last_ms, t); }
    if (bad_block < 0) { return SRLA_APIRESULT_OK; }
    /* the blocks of the bad block's group that precede it were decoded and are delivered like the reference does */
    const uint32_t from = smp_begin(published), to = blocks[(size_t)bad_block].sample_offset;
    for (uint32_t ch = 0; ch < nch && to > from; ch++) { std::memcpy(buffer[ch] + from, (const int32_t *)c->h_out.p + stride * ch + from, sizeof(int32_t) * (to - from)); }
    const uint32_t code = st[bad_block];
    return (code <= (uint32_t)SRLA_APIRESULT_NG) ? (SRLAApiResult)code : SRLA_APIRESULT_NG;
}

/* decode `blocks` (already walked) of the stream at `data` into the caller's channel pointers */
SRLAApiResult decoder_run(struct SRLADecoder *d, const uint8_t *data, uint64_t data_bytes, const std::vector<DecBlock> &blocks,
                          int32_t **buffer, uint32_t total_samples)
{
    DecoderCtx *c = d->ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    const uint32_t nch = d->header.num_channels;
    const uint64_t stride = round_up_u32(total_samples, 16);
    const size_t nb = blocks.size();
    if (nb == 0) { return SRLA_APIRESULT_OK; }
    if (nb >= 256 && c->host_threads > 0) {
        /* the pipelined path page-locks staging memory as large as the stream and its PCM (~0.7 ms per MB, once per
         * handle): worth it for a handle that keeps decoding, not for a one-shot tool */
        c->long_calls++;
        if (c->pipeline == 1 || (c->pipeline < 0 && c->long_calls >= 2)) { return decoder_run_pipelined(d, data, data_bytes, blocks, buffer, total_samples); }
    }
    if (!c->data.reserve(data_bytes + 16) || !c->out.reserve(sizeof(int32_t) * stride * nch) || !c->blocks.reserve(sizeof(DecBlock) * nb)
        || !c->status.reserve(sizeof(uint32_t) * nb) || !c->side.reserve(sizeof(DecSide) * nb) || !c->side_ch.reserve(sizeof(DecSideChannel) * nb * nch) || !c->h_blocks.reserve(sizeof(DecBlock) * nb) || !c->h_status.reserve(sizeof(uint32_t) * nb)) { return SRLA_APIRESULT_NG; }
    std::memcpy(c->h_blocks.p, blocks.data(), sizeof(DecBlock) * nb);
    if (cudaMemcpyAsync(c->blocks.p, c->h_blocks.p, sizeof(DecBlock) * nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess
        || cudaMemcpyAsync(c->data.p, data, data_bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess
        || cudaMemsetAsync((uint8_t *)c->data.p + data_bytes, 0, 16, c->stream) != cudaSuccess
        || cudaMemsetAsync(c->status.p, 0xff, sizeof(uint32_t) * nb, c->stream) != cudaSuccess
        || cudaMemsetAsync(c->side.p, 0, sizeof(DecSide) * nb, c->stream) != cudaSuccess
        || cudaMemsetAsync(c->side_ch.p, 0, sizeof(DecSideChannel) * nb * nch, c->stream) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    DecParams p;
    p.data = (const uint8_t *)c->data.p; p.blocks = (const DecBlock *)c->blocks.p; p.status = (uint32_t *)c->status.p;
    p.out = (int32_t *)c->out.p; p.stride = stride;
    p.nch = nch; p.bps = d->header.bits_per_sample; p.lshift = d->header.offset_lshift; p.check = (d->config.check_checksum == 1) ? 1u : 0u;
    p.tree = (const uint16_t *)c->tree.p;
    p.side = (DecSide *)c->side.p; p.side_ch = (DecSideChannel *)c->side_ch.p; p.num_blocks = (uint32_t)nb; p.pad = 0;
    cudaEventRecord(c->ev0, c->stream);
    const unsigned pl = parse_lanes_for(c, nb);
    decode_parse_kernel<<<(unsigned)((nb + pl - 1) / pl), pl, 0, c->stream>>>(p);
    decode_blocks_kernel<<<(unsigned)nb, 32u * std::max(1u, nch), nch * (sizeof(DecChannel) + sizeof(int32_t) * kDecStage), c->stream>>>(p);
    cudaEventRecord(c->ev1, c->stream);
    if (cudaGetLastError() != cudaSuccess) { return SRLA_APIRESULT_NG; }
    if (cudaMemcpyAsync(c->h_status.p, c->status.p, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess
        || cudaStreamSynchronize(c->stream) != cudaSuccess) {
        std::fprintf(stderr, "[srla_b200] decode failed on the device: %s\n", cudaGetErrorString(cudaGetLastError()));
        return SRLA_APIRESULT_NG;
    }
    cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
    /* the reference stops at the first bad block and leaves the earlier ones decoded (srla_decoder.c:780-786) */
    const uint32_t *st = (const uint32_t *)c->h_status.p;
    uint32_t good_samples = total_samples; SRLAApiResult rc = SRLA_APIRESULT_OK;
    for (size_t i = 0; i < nb; i++) { if (st[i] != 0u) { rc = (st[i] <= (uint32_t)SRLA_APIRESULT_NG) ? (SRLAApiResult)st[i] : SRLA_APIRESULT_NG; good_samples = blocks[i].sample_offset; break; } }
    for (uint32_t ch = 0; ch < nch && good_samples > 0; ch++) {
        if (cudaMemcpy(buffer[ch], (const int32_t *)c->out.p + stride * ch, sizeof(int32_t) * good_samples, cudaMemcpyDeviceToHost) != cudaSuccess) { return SRLA_APIRESULT_NG; }
    }
    return rc;
}

} // namespace

extern "C" {

SRLAApiResult SRLADecoder_DecodeHeader(const uint8_t *data, uint32_t data_size, struct SRLAHeader *header)
{
    if (data == NULL || header == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (data_size < SRLA_HEADER_SIZE) { return SRLA_APIRESULT_INSUFFICIENT_DATA; }
    if (data[0] != '1' || data[1] != '2' || data[2] != '4' || data[3] != '9') { return SRLA_APIRESULT_INVALID_FORMAT; }
    auto be32 = [&](int at) { return ((uint32_t)data[at] << 24) | ((uint32_t)data[at + 1] << 16) | ((uint32_t)data[at + 2] << 8) | data[at + 3]; };
    auto be16 = [&](int at) { return (uint16_t)(((uint32_t)data[at] << 8) | data[at + 1]); };
    struct SRLAHeader h;
    h.format_version = be32(4); h.codec_version = be32(8); h.num_channels = be16(12); h.num_samples = be32(14);
    h.sampling_rate = be32(18); h.bits_per_sample = be16(22); h.offset_lshift = data[24];
    h.max_num_samples_per_block = be32(25); h.preset = data[29];
    *header = h;
    return SRLA_APIRESULT_OK;
}

int32_t SRLADecoder_CalculateWorkSize(const struct SRLADecoderConfig *config)
{
    if (config == NULL || config->max_num_channels == 0) { return -1; }
    return (int32_t)(sizeof(struct SRLADecoder) + 64);
}

struct SRLADecoder *SRLADecoder_Create(const struct SRLADecoderConfig *config, void *work, int32_t work_size)
{
    uint8_t own = 0;
    if (work == NULL && work_size == 0) {
        if ((work_size = SRLADecoder_CalculateWorkSize(config)) < 0) { return NULL; }
        work = std::malloc((size_t)work_size);
        own = 1;
    }
    if (config == NULL || work == NULL || work_size < SRLADecoder_CalculateWorkSize(config) || config->max_num_channels == 0
        || config->max_num_channels > SRLA_MAX_NUM_CHANNELS) {
        if (own) { std::free(work); }
        return NULL;
    }
    struct SRLADecoder *d = (struct SRLADecoder *)(((uintptr_t)work + 15u) & ~(uintptr_t)15u);
    std::memset(d, 0, sizeof(*d));
    d->magic = kDecoderMagic; d->config = *config; d->alloced_by_own = own; d->work = work;
    d->ctx = new (std::nothrow) DecoderCtx();
    if (d->ctx == nullptr || !decoder_ctx_init(d->ctx)) {
        if (d->ctx) { decoder_ctx_destroy(d->ctx); delete d->ctx; }
        d->magic = 0;
        if (own) { std::free(work); }
        return NULL;
    }
    return d;
}

void SRLADecoder_Destroy(struct SRLADecoder *decoder)
{
    if (decoder == NULL || decoder->magic != kDecoderMagic) { return; }
    decoder_ctx_destroy(decoder->ctx);
    delete decoder->ctx;
    decoder->magic = 0;
    if (decoder->alloced_by_own) { std::free(decoder->work); }
}

SRLAApiResult SRLADecoder_SetHeader(struct SRLADecoder *decoder, const struct SRLAHeader *header)
{
    if (decoder == NULL || header == NULL || decoder->magic != kDecoderMagic) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (decoder_header_valid(header) != SRLA_APIRESULT_OK) { return SRLA_APIRESULT_INVALID_FORMAT; }
    if (decoder->config.max_num_channels < header->num_channels) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    if (decoder->config.max_num_parameters < kPresetMaxOrder[header->preset]) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    if (header->bits_per_sample != 8 && header->bits_per_sample != 16 && header->bits_per_sample != 24) { return SRLA_APIRESULT_INVALID_FORMAT; }
    decoder->header = *header;
    decoder->set_header = 1;
    return SRLA_APIRESULT_OK;
}

SRLAApiResult SRLADecoder_DecodeBlock(
    struct SRLADecoder *decoder, const uint8_t *data, uint32_t data_size,
    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples, uint32_t *decode_size, uint32_t *num_decode_samples)
{
    if (decoder == NULL || data == NULL || buffer == NULL || decode_size == NULL || num_decode_samples == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (decoder->magic != kDecoderMagic) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (!decoder->set_header) { return SRLA_APIRESULT_PARAMETER_NOT_SET; }
    if (buffer_num_channels < decoder->header.num_channels) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    if (data_size < 11u) { return SRLA_APIRESULT_INSUFFICIENT_DATA; }
    if (data[0] != 0xFF || data[1] != 0xFF) { return SRLA_APIRESULT_INVALID_FORMAT; }
    const uint32_t size = ((uint32_t)data[2] << 24) | ((uint32_t)data[3] << 16) | ((uint32_t)data[4] << 8) | data[5];
    if ((uint64_t)size + 6u > data_size) { return SRLA_APIRESULT_INSUFFICIENT_DATA; }
    const uint32_t n = ((uint32_t)data[9] << 8) | data[10];
    std::vector<DecBlock> one(1);
    one[0].offset = 0; one[0].bytes = size + 6u; one[0].sample_offset = 0; one[0].nsmpl = n; one[0].pad = 0;
    if (size < 5u || n == 0u) { return SRLA_APIRESULT_INVALID_FORMAT; }
    if (n > buffer_num_samples) {
        /* a bad checksum outranks the capacity error (srla_decoder.c:683-700) */
        if (decoder->config.check_checksum == 1 && host::fletcher16(data + 8, size - 2u) != (uint16_t)(((uint32_t)data[6] << 8) | data[7])) { return SRLA_APIRESULT_DETECT_DATA_CORRUPTION; }
        return SRLA_APIRESULT_INSUFFICIENT_BUFFER;
    }
    const SRLAApiResult rc = decoder_run(decoder, data, size + 6u, one, buffer, n);
    if (rc != SRLA_APIRESULT_OK) { return rc; }
    *decode_size = size + 6u;
    *num_decode_samples = n;
    return SRLA_APIRESULT_OK;
}

SRLAApiResult SRLADecoder_DecodeWhole(
    struct SRLADecoder *decoder, const uint8_t *data, uint32_t data_size,
    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples)
{
    if (decoder == NULL || data == NULL || buffer == NULL) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    if (decoder->magic != kDecoderMagic) { return SRLA_APIRESULT_INVALID_ARGUMENT; }
    struct SRLAHeader h;
    SRLAApiResult rc = SRLADecoder_DecodeHeader(data, data_size, &h);
    if (rc != SRLA_APIRESULT_OK) { return rc; }
    if ((rc = SRLADecoder_SetHeader(decoder, &h)) != SRLA_APIRESULT_OK) { return rc; }
    if (buffer_num_channels < h.num_channels || buffer_num_samples < h.num_samples) { return SRLA_APIRESULT_INSUFFICIENT_BUFFER; }
    /* host walk over the block headers (srla_decoder.c:768-795): the size fields chain the blocks */
    std::vector<DecBlock> blocks;
    uint32_t progress = 0; uint64_t at = SRLA_HEADER_SIZE;
    SRLAApiResult walk = SRLA_APIRESULT_OK;
    while (progress < h.num_samples && at < data_size) {
        const uint64_t left = data_size - at;
        if (left < 11u) { walk = SRLA_APIRESULT_INSUFFICIENT_DATA; break; }
        const uint8_t *b = data + at;
        if (b[0] != 0xFF || b[1] != 0xFF) { walk = SRLA_APIRESULT_INVALID_FORMAT; break; }
        const uint32_t size = ((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5];
        if ((uint64_t)size + 6u > left) { walk = SRLA_APIRESULT_INSUFFICIENT_DATA; break; }
        if (size < 5u) { walk = SRLA_APIRESULT_INVALID_FORMAT; break; }
        const uint32_t n = ((uint32_t)b[9] << 8) | b[10];
        if (n == 0u) { walk = SRLA_APIRESULT_INVALID_FORMAT; break; }       /* the reference's encoder never writes one; the loop would not advance */
        if (n > buffer_num_samples - progress) {
            /* a bad checksum outranks the capacity error (srla_decoder.c:683-700) */
            const bool corrupt = decoder->config.check_checksum == 1 && host::fletcher16(b + 8, size - 2u) != (uint16_t)(((uint32_t)b[6] << 8) | b[7]);
            walk = corrupt ? SRLA_APIRESULT_DETECT_DATA_CORRUPTION : SRLA_APIRESULT_INSUFFICIENT_BUFFER;
            break;
        }
        DecBlock blk; blk.offset = at; blk.bytes = size + 6u; blk.sample_offset = progress; blk.nsmpl = n; blk.pad = 0;
        blocks.push_back(blk);
        at += (uint64_t)size + 6u; progress += n;
    }
    if (progress == 0u) { return walk; }                                   /* no block to decode: nothing is copied back */
    rc = decoder_run(decoder, data, at, blocks, buffer, progress);
    return (rc != SRLA_APIRESULT_OK) ? rc : walk;
}

float SRLAB200_DecoderKernelMs(const struct SRLADecoder *decoder)
{
    return (decoder && decoder->magic == kDecoderMagic) ? decoder->ctx->last_ms : -1.0f;
}

} /* extern "C" */

#endif
