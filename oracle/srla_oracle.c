/*
 * srla_oracle.c -- CPU restatement of the SRLA encode path (format 10 / codec 18) and, at the end of the
 * file, of the decoder (pinned by tests/test_oracle.py against the reference decoder and the committed streams).
 *
 * TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared with.  It is never the thing
 * shipped or measured (except as bench.py's "port" CPU baseline).  See srla_oracle.h.
 *
 * Written from the behaviour of the reference (paths relative to /root/reference); every stage
 * cites the file:line it follows.  Parity status: PINNED -- tests/test_oracle.py compares
 * whole .srl streams byte for byte with the compiled reference (oracle/_ref/libsrla_ref.so) and
 * with the committed reference-generated fixtures in tests/golden/.
 *
 * Stale scratch memory.  Three corners of the reference read what EARLIER calls left in the LPC calculator's
 * scratch (lpc.c:211-216 `buffer`, lpc.c:188 `auto_corr`), so its output there depends on the handle's history:
 *  - odd block length: the Welch window never writes the middle sample of `buffer` (lpc.c:260-264); it keeps the
 *    previous call's inverse-transform output at that index;
 *  - LTP with an FFT size N < 263: lags N..262 are copied from beyond the transformed region (lpc.c:371-373);
 *  - the LTP pitch search may read lags 263/264, which nothing ever writes (lpc.c:1497-1513).
 * This restatement keeps the same persistent scratch (lpc_scratch below) and walks the calls in the reference's
 * order, so it reproduces all three for a handle that starts from zeroed memory -- which is what a fresh `srla`
 * CLI process has (its work area is one large malloc, i.e. fresh zero pages) and what so_encode_whole() models by
 * resetting the scratch first.  so_reset_state() does the same for the single-block entry points.
 * Not modelled: an SVR run (num_svr_filter_learning_iteration > 0) also parks its residual in `buffer`
 * (lpc.c:1049), and a max block size below 263 with LTP lets the reference read past `buffer`.
 *
 * Floating point: compile with -ffp-contract=off (oracle/Makefile does); the reference is built as
 * ISO C90, i.e. without FMA contraction, and byte-identical output needs the same roundings.
 */
#include "srla_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/srla_format_tables.h"

/* ------------------------------------------------------------------------------------------------
 * format constants (libs/srla_internal/include/srla_internal.h:13-35, include/srla.h:6-24)
 * ---------------------------------------------------------------------------------------------- */
enum {
    FMT_VERSION = 10, CODEC_VERSION = 18, FILE_HEADER_BYTES = 30, BLOCK_HEADER_BYTES = 11,
    PRE_SHIFT = 4, COEF_BITS = 8, RSHIFT_BITS = 4, ORDER_BITS = 8,
    LTP_COEF_BITS = 6, LTP_PERIOD_BITS = 8, LTP_MIN_PERIOD = 8, LTP_MAX_PERIOD = 8 + 256 - 2,
    BLOCK_COMPRESS = 0, BLOCK_SILENT = 1, BLOCK_RAW = 2,
    CODE_RICE = 0, CODE_RECURSIVE_RICE = 1, CODE_ALLZERO = 2,
    LOG2_MAX_PARTS = 10, RICE_PARAM_BITS = 5
};
static const uint32_t PRESET_MAX_ORDER[7] = { 0, 8, 16, 32, 64, 128, 255 }; /* srla_internal.c:30-38 */
#define RIDGE 1e-5                                                         /* srla_internal.h:23 */

static uint32_t zigzag(int32_t v) { return (uint32_t)((-(v < 0)) ^ (int32_t)((uint32_t)v << 1)); } /* srla_utility.h:31 */
static double round_half_away(double d) { return (d >= 0.0) ? floor(d + 0.5) : -floor(-d + 0.5); } /* srla_utility.c:22 */
static double log2_via_ln(double d) { return log(d) * 1.4426950408889634; }                       /* srla_utility.c:28 */
static uint32_t floor_log2(uint32_t v) { uint32_t r = 0; while (v >>= 1) { r++; } return r; }
static uint32_t ceil_pow2(uint32_t v) { uint32_t p = 1; while (p < v) { p <<= 1; } return p; }

/* ------------------------------------------------------------------------------------------------
 * Fletcher-16 (srla_utility.c:36-60; both running sums are plain residues mod 255)
 * ---------------------------------------------------------------------------------------------- */
uint16_t so_fletcher16(const uint8_t *data, size_t size)
{
    uint32_t lo = 0, hi = 0;
    size_t i;
    for (i = 0; i < size; i++) {
        lo = (lo + data[i]) % 255u;
        hi = (hi + lo) % 255u;
    }
    return (uint16_t)((hi << 8) | lo);
}

/* ------------------------------------------------------------------------------------------------
 * static Huffman code construction (static_huffman.c:28-132)
 * Repeatedly merges the two live nodes that are smallest under (count, index) order; the smaller
 * becomes the 0-branch.  Zero counts are bumped to one.  Codes are read root -> leaf.
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t FREQ_PLAIN[256] = SRLA_FMT_COEF_SYMBOL_FREQ_INIT;
static const uint32_t FREQ_SUMMED[256] = SRLA_FMT_SUMMED_COEF_SYMBOL_FREQ_INIT;

typedef struct { uint32_t code[256]; uint8_t len[256]; } huff_table;

static void huff_assign(const uint32_t (*kids)[2], uint32_t nsym, uint32_t node, uint32_t code, uint8_t len, huff_table *t)
{
    if (node < nsym) { t->code[node] = code; t->len[node] = len; return; }
    huff_assign(kids, nsym, kids[node][0], (code << 1) | 0u, (uint8_t)(len + 1), t);
    huff_assign(kids, nsym, kids[node][1], (code << 1) | 1u, (uint8_t)(len + 1), t);
}

static void huff_build(const uint32_t *counts, uint32_t nsym, huff_table *t)
{
    uint32_t weight[512]; uint8_t live[512]; uint32_t kids[512][2];
    uint32_t total = nsym, i;
    memset(t, 0, sizeof(*t));
    for (i = 0; i < nsym; i++) { weight[i] = counts[i] ? counts[i] : 1u; live[i] = 1; }
    for (;;) {
        int a = -1, b = -1;
        for (i = 0; i < total; i++) {
            if (!live[i]) { continue; }
            if (a < 0 || weight[i] < weight[a]) { b = a; a = (int)i; }
            else if (b < 0 || weight[i] < weight[b]) { b = (int)i; }
        }
        if (b < 0) { huff_assign((const uint32_t (*)[2])kids, nsym, (uint32_t)a, 0, 0, t); return; }
        weight[total] = weight[a] + weight[b];
        live[total] = 1; live[a] = live[b] = 0;
        kids[total][0] = (uint32_t)a; kids[total][1] = (uint32_t)b;
        total++;
    }
}

static const huff_table *format_huffman(int summed)
{
    static huff_table tables[2]; static int ready = 0;
    if (!ready) { huff_build(FREQ_PLAIN, 256, &tables[0]); huff_build(FREQ_SUMMED, 256, &tables[1]); ready = 1; }
    return &tables[summed ? 1 : 0];
}

void so_huffman_codes(int which, const uint32_t *counts, uint32_t num_symbols, uint32_t *codes, uint8_t *lens)
{
    huff_table t; uint32_t i;
    if (which == 0) { counts = FREQ_PLAIN; num_symbols = 256; }
    if (which == 1) { counts = FREQ_SUMMED; num_symbols = 256; }
    huff_build(counts, num_symbols, &t);
    for (i = 0; i < num_symbols; i++) { codes[i] = t.code[i]; lens[i] = t.len[i]; }
}

/* ------------------------------------------------------------------------------------------------
 * MSB-first bit writer (bit_stream.h:245-307, 400-437).  Target memory must be zeroed.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint8_t *mem; uint64_t bit; } bitsink;

static void put_bits(bitsink *s, uint32_t value, uint32_t nbits)
{
    while (nbits) {
        nbits--;
        if ((value >> nbits) & 1u) { s->mem[s->bit >> 3] |= (uint8_t)(0x80u >> (s->bit & 7u)); }
        s->bit++;
    }
}
static void put_zero_run(bitsink *s, uint32_t run) { s->bit += run; put_bits(s, 1u, 1u); } /* `run` zeros then a one */
static uint32_t sink_bytes(const bitsink *s) { return (uint32_t)((s->bit + 7u) >> 3); }

static void store_be(uint8_t *p, uint32_t v, int nbytes) { int i; for (i = 0; i < nbytes; i++) { p[i] = (uint8_t)(v >> (8 * (nbytes - 1 - i))); } }

/* ------------------------------------------------------------------------------------------------
 * file header (srla_encoder.c:85-165) and trailing-zero shift (srla_utility.c:177-203)
 * ---------------------------------------------------------------------------------------------- */
int so_encode_header(const so_params *p, uint32_t num_samples, uint8_t *out, uint32_t cap)
{
    if (!p || !out) { return SO_INVALID_ARGUMENT; }
    if (cap < FILE_HEADER_BYTES) { return SO_INSUFFICIENT_BUFFER; }
    if (!p->num_channels || !num_samples || !p->sampling_rate || !p->bits_per_sample
        || p->offset_lshift >= 32 || !p->max_block || p->preset >= 7) { return SO_INVALID_FORMAT; }
    out[0] = '1'; out[1] = '2'; out[2] = '4'; out[3] = '9';
    store_be(out + 4, FMT_VERSION, 4);
    store_be(out + 8, CODEC_VERSION, 4);
    store_be(out + 12, p->num_channels, 2);
    store_be(out + 14, num_samples, 4);
    store_be(out + 18, p->sampling_rate, 4);
    store_be(out + 22, p->bits_per_sample, 2);
    out[24] = (uint8_t)p->offset_lshift;
    store_be(out + 25, p->max_block, 4);
    out[29] = (uint8_t)p->preset;
    return SO_OK;
}

uint32_t so_offset_lshift(const int32_t *const *pcm, uint32_t num_channels, uint32_t num_samples)
{
    uint32_t used = 0, c, i, shift = 0;
    for (c = 0; c < num_channels; c++) { for (i = 0; i < num_samples; i++) { used |= (uint32_t)pcm[c][i]; } }
    if (!used) { return 0; }
    while (!((used >> shift) & 1u)) { shift++; }
    return shift;
}

/* ------------------------------------------------------------------------------------------------
 * FFT exactly as the reference evaluates it (libs/fft/src/fft.c)
 *   complex radix-4 Stockham, twiddles by the recurrence w <- w * wdelta      fft.c:71-128
 *   real-FFT wrapper (split/merge with its own twiddle recurrence)            fft.c:147-198
 * ---------------------------------------------------------------------------------------------- */
#define FFT_PI 3.14159265358979323846
typedef struct { double re, im; } cpx;
static cpx c_add(cpx a, cpx b) { cpx r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
static cpx c_sub(cpx a, cpx b) { cpx r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
static cpx c_mul(cpx a, cpx b) { cpx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }

static void complex_fft(int n, int flag, cpx *x, cpx *y)
{
    cpx *src = x, *in = x, *out = y, *t;
    int stride = 1, p, q;
    const cpx jrot = { 0.0, (double)(-flag) };
    while (n > 2) {
        const int quarter = n >> 2, half = n >> 1;
        const double theta = 2.0 * FFT_PI / n;
        cpx step, w1;
        step.re = cos(theta); step.im = flag * sin(theta);
        w1.re = 1.0; w1.im = 0.0;
        for (p = 0; p < quarter; p++) {
            const cpx w2 = c_mul(w1, w1);
            const cpx w3 = c_mul(w1, w2);
            for (q = 0; q < stride; q++) {
                const cpx a = in[q + stride * p];
                const cpx b = in[q + stride * (p + quarter)];
                const cpx c = in[q + stride * (p + half)];
                const cpx d = in[q + stride * (p + half + quarter)];
                const cpx ac_sum = c_add(a, c), ac_dif = c_sub(a, c), bd_sum = c_add(b, d);
                const cpx bd_rot = c_mul(jrot, c_sub(b, d));
                out[q + stride * (4 * p + 0)] = c_add(ac_sum, bd_sum);
                out[q + stride * (4 * p + 1)] = c_mul(w1, c_sub(ac_dif, bd_rot));
                out[q + stride * (4 * p + 2)] = c_mul(w2, c_sub(ac_sum, bd_sum));
                out[q + stride * (4 * p + 3)] = c_mul(w3, c_add(ac_dif, bd_rot));
            }
            w1 = c_mul(w1, step);
        }
        n >>= 2; stride <<= 2;
        t = in; in = out; out = t;
    }
    if (n == 2) {
        for (q = 0; q < stride; q++) {
            const cpx a = in[q], b = in[q + stride];
            out[q] = c_add(a, b);
            out[q + stride] = c_sub(a, b);
        }
        stride <<= 1;
        t = in; in = out; out = t;
    }
    if (in != src) { memcpy(src, in, sizeof(cpx) * (size_t)stride); }
}

static void real_fft_work(int n, int flag, double *x, double *work)
{
    const double theta = flag * 2.0 * FFT_PI / n;
    const double dsin = sin(theta);
    const double dcosm1 = cos(theta) - 1.0;
    const double c2 = flag * 0.5;
    double wr, wi;
    int i;
    if (flag == -1) { complex_fft(n >> 1, -1, (cpx *)x, (cpx *)work); }
    wr = 1.0 + dcosm1; wi = dsin;
    for (i = 1; i <= (n >> 2); i++) {
        const int lo_r = 2 * i, lo_i = lo_r + 1, hi_r = n - lo_r, hi_i = hi_r + 1;
        const double h1r = 0.5 * (x[lo_r] + x[hi_r]);
        const double h1i = 0.5 * (x[lo_i] - x[hi_i]);
        const double h2r = -c2 * (x[lo_i] + x[hi_i]);
        const double h2i = c2 * (x[lo_r] - x[hi_r]);
        double keep;
        x[lo_r] = h1r + (wr * h2r) - (wi * h2i);
        x[lo_i] = h1i + (wr * h2i) + (wi * h2r);
        x[hi_r] = h1r - (wr * h2r) + (wi * h2i);   /* for i == n/4 these overwrite the two stores above */
        x[hi_i] = -h1i + (wr * h2i) + (wi * h2r);
        keep = wr;
        wr += keep * dcosm1 - wi * dsin;
        wi += wi * dcosm1 + keep * dsin;
    }
    {
        const double dc = x[0];
        if (flag == -1) { x[0] = dc + x[1]; x[1] = dc - x[1]; }
        else { x[0] = 0.5 * (dc + x[1]); x[1] = 0.5 * (dc - x[1]); complex_fft(n >> 1, 1, (cpx *)x, (cpx *)work); }
    }
}

void so_real_fft(int n, int flag, double *x)
{
    double *work = (double *)malloc(sizeof(double) * (size_t)n);
    real_fft_work(n, flag, x, work);
    free(work);
}

/* ------------------------------------------------------------------------------------------------
 * Welch window (lpc.c:252-266) + autocorrelation through the FFT (lpc.c:330-376):
 *   r[k] = (2/n) * IFFT(|FFT(xw, zero padded to N)|^2)[k],  N = next power of two >= n
 * which is (N/n) x the circular autocorrelation over N (no padding when n is a power of two).
 * ---------------------------------------------------------------------------------------------- */
/* the reference calculator's persistent scratch: `buffer` (lpc.c:211-213) survives from call to call */
static struct { double *buffer; uint32_t cap; } lpc_scratch;
static double *lpc_scratch_reserve(uint32_t want)
{
    if (want > lpc_scratch.cap) {
        uint32_t cap = lpc_scratch.cap ? lpc_scratch.cap : 1024u;
        double *grown;
        while (cap < want) { cap <<= 1; }
        grown = (double *)calloc((size_t)cap, sizeof(double));       /* memory the reference never touched reads as zero */
        if (lpc_scratch.buffer) { memcpy(grown, lpc_scratch.buffer, sizeof(double) * lpc_scratch.cap); free(lpc_scratch.buffer); }
        lpc_scratch.buffer = grown; lpc_scratch.cap = cap;
    }
    return lpc_scratch.buffer;
}
void so_reset_state(void) { if (lpc_scratch.buffer) { memset(lpc_scratch.buffer, 0, sizeof(double) * lpc_scratch.cap); } }

static void welch_autocorr(const double *x, uint32_t n, double *r, uint32_t nlags)
{
    const uint32_t N = ceil_pow2(n);
    double *buf = lpc_scratch_reserve(((N > nlags) ? N : nlags) + 2u);
    double *work = (double *)calloc((size_t)N + 2, sizeof(double));
    const double divisor = 4.0 * pow((double)(n - 1), -2.0);
    const double scale = 2.0 / n;
    uint32_t i;
    /* lpc.c:260-264: both window halves, i < n / 2 -- for odd n the middle sample keeps what the previous call left */
    for (i = 0; i < (n >> 1); i++) {
        const double w = divisor * i * (n - 1 - i);
        buf[i] = x[i] * w;
        buf[n - 1 - i] = x[n - 1 - i] * w;
    }
    for (i = n; i < N; i++) { buf[i] = 0.0; }                                   /* lpc.c:349-351 */
    if (N >= 2) {
        real_fft_work((int)N, -1, buf, work);
        buf[0] *= buf[0];
        buf[1] *= buf[1];
        for (i = 2; i < N; i += 2) { const double a = buf[i], b = buf[i + 1]; buf[i] = a * a + b * b; buf[i + 1] = 0.0; }
        real_fft_work((int)N, 1, buf, work);
    }
    for (i = 0; i < nlags; i++) { r[i] = buf[i] * scale; }                      /* lpc.c:371-373: also beyond N */
    free(work);
}

void so_autocorr(const double *x, uint32_t n, double *r, uint32_t max_lag) { welch_autocorr(x, n, r, max_lag + 1); }

/* ------------------------------------------------------------------------------------------------
 * Levinson-Durbin for all orders 1..P (lpc.c:379-441), sequential dot product for the reflection
 * coefficient; rows[k] = coefficient vector of order k+1 with rows[k][0] == 1.
 * ---------------------------------------------------------------------------------------------- */
static void levinson(const double *r, uint32_t P, double (*rows)[SO_MAX_ORDER + 3], double *err)
{
    uint32_t k, i;
    if (fabs(r[0]) < FLT_EPSILON) {
        for (i = 0; i <= P; i++) { err[i] = r[0]; }
        for (k = 0; k < P; k++) { for (i = 0; i < P + 2; i++) { rows[k][i] = 0.0; } }
        return;
    }
    rows[0][0] = 1.0;
    err[0] = r[0];
    rows[0][1] = -r[1] / r[0];
    rows[0][2] = 0.0;
    err[1] = err[0] + r[1] * rows[0][1];
    for (k = 1; k < P; k++) {
        const double *prev = rows[k - 1];
        double refl = 0.0;
        for (i = 0; i <= k; i++) { refl += prev[i] * r[k + 1 - i]; }
        refl /= -err[k];
        err[k + 1] = err[k] * (1.0 - refl * refl);
        for (i = 0; i < k + 2; i++) { rows[k][i] = prev[i] + refl * prev[k + 1 - i]; }
        rows[k][k + 2] = 0.0;
    }
}

/* Welch-window energy compensation (lpc.c:275-290) */
static double welch_energy_gain(uint32_t n_samples)
{
    const double n = n_samples - 1;
    return (15 * (n - 1) * (n - 1) * (n - 1)) / (8 * n * (n - 2) * (n * n - 2 * n + 2));
}

/* entropy of the geometric distribution whose mean is `mean_abs` full-scale units (srla_encoder.c:873-885) */
static double geometric_entropy(double mean_abs, uint32_t bps)
{
    const double int_mean = mean_abs * (1 << (bps - 1));
    const double rho = 1.0 / (1.0 + int_mean);
    const double inv = 1.0 - rho;
    if (mean_abs < 1e-16) { return 0.0; }
    return -(inv * log2_via_ln(inv) + rho * log2_via_ln(rho)) / rho;
}

/* first order minimising estimated bits (srla_encoder.c:934-957) */
static uint32_t choose_order(const double *err, uint32_t P, uint32_t n, uint32_t bps)
{
    double best = FLT_MAX; uint32_t arg = 0, k;
    for (k = 1; k <= P; k++) {
        const double mean_abs = 2.0 * sqrt(err[k] / 2.0);
        double bits = geometric_entropy(mean_abs, bps) * n;
        bits += COEF_BITS * k;
        if (best > bits) { best = bits; arg = k; }
    }
    return arg;
}

/* 8-bit quantisation with error feedback from the tail (lpc.c:1341-1405) */
static void quantise_lpc(const double *a, uint32_t order, int32_t *q, uint32_t *rshift_out)
{
    const int32_t limit = 1 << (COEF_BITS - 1);
    double peak = 0.0, carry = 0.0;
    int32_t i, exponent; uint32_t rshift;
    for (i = 0; i < (int32_t)order; i++) { if (peak < fabs(a[i])) { peak = fabs(a[i]); } }
    if (peak <= pow(2.0, -(COEF_BITS - 1))) {
        *rshift_out = COEF_BITS;
        memset(q, 0, sizeof(int32_t) * order);
        return;
    }
    (void)frexp(peak, &exponent);
    rshift = (uint32_t)((COEF_BITS - 1) - exponent);
    if (rshift >= (1u << RSHIFT_BITS)) { rshift = (1u << RSHIFT_BITS) - 1; }
    for (i = (int32_t)order - 1; i >= 0; i--) {
        int32_t v;
        carry += a[i] * pow(2.0, (double)rshift);
        v = (int32_t)round_half_away(carry);
        if (v >= limit) { v = limit - 1; } else if (v < -limit) { v = -limit; }
        carry -= v;
        q[i] = v;
    }
    *rshift_out = rshift;
}

/* ------------------------------------------------------------------------------------------------
 * integer filters: pre-emphasis (srla_utility.c:214-257, 342-358), FIR residual
 * (srla_lpc_predict.c:236-264), LTP residual (srla_lpc_predict.c:267-294).  int32 wraps.
 * ---------------------------------------------------------------------------------------------- */
static int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static int32_t asr(int32_t v, uint32_t s) { return (s >= 32) ? (v < 0 ? -1 : 0) : (int32_t)(v >> s); }

static int32_t preemphasis_coef(const int32_t *x, uint32_t n)
{
    double r0 = 0.0, r1 = 0.0; uint32_t i; int32_t c;
    for (i = 0; i + 1 < n; i++) { const double a = x[i], b = x[i + 1]; r0 += a * a; r1 += a * b; }
    { const double a = x[n - 1]; r0 += a * a; }
    if (r0 < 1e-6) { return 0; }
    c = (int32_t)round_half_away((r1 / r0) * pow(2.0, PRE_SHIFT));
    if (c < -(1 << PRE_SHIFT)) { c = -(1 << PRE_SHIFT); }
    if (c > (1 << PRE_SHIFT) - 1) { c = (1 << PRE_SHIFT) - 1; }
    return c;
}

static void preemphasis_apply(int32_t *x, uint32_t n, int32_t coef)
{
    int32_t prev = x[0]; uint32_t i;      /* filter memory seeded with the first sample (srla_encoder.c:998-1003) */
    for (i = 0; i < n; i++) {
        const int32_t cur = x[i];
        x[i] = wrap_sub(cur, asr(wrap_mul(prev, coef), PRE_SHIFT));
        prev = cur;
    }
}

static void fir_residual(const int32_t *x, uint32_t n, const int32_t *c, uint32_t order, uint32_t rshift, int32_t *res)
{
    /* rshift == 0 never occurs for |coef| <= 127 quantised from max >= 2^-7; x86 would give half = 1<<31 */
    const int32_t half = (rshift > 0) ? (int32_t)(1u << (rshift - 1)) : (int32_t)0x80000000u;
    uint32_t i, j;
    res[0] = x[0];
    for (i = 1; i < order && i < n; i++) { res[i] = wrap_sub(x[i], x[i - 1]); }
    for (i = order; i < n; i++) {
        int32_t acc = half;
        for (j = 0; j < order; j++) { acc = wrap_add(acc, wrap_mul(c[j], x[i - order + j])); }
        res[i] = wrap_add(x[i], asr(acc, rshift));
    }
}

static void ltp_residual(const int32_t *x, uint32_t n, const int32_t *c, uint32_t order, uint32_t period, int32_t *res)
{
    const uint32_t half_order = order >> 1, shift = LTP_COEF_BITS - 1;
    uint32_t i, j;
    memcpy(res, x, sizeof(int32_t) * n);
    for (i = period + half_order + 1; i < n; i++) {
        int32_t acc = 1 << (shift - 1);
        for (j = 0; j < order; j++) { acc = wrap_add(acc, wrap_mul(c[j], x[i - period - half_order + j])); }
        res[i] = wrap_sub(x[i], asr(acc, shift));
    }
}

/* ------------------------------------------------------------------------------------------------
 * long-term (pitch) predictor: pitch pick (lpc.c:1473-1555), 3-tap normal equations by Cholesky
 * (lpc.c:573-631, 1558-1649), 6-bit quantisation (srla_encoder.c:1032-1047).
 * returns 0 ok (period may be 0 = no pitch), SO_NG when the reference would fail the whole encode.
 * ---------------------------------------------------------------------------------------------- */
static int detect_pitch(const double *r /* lags 0..LTP_MAX_PERIOD+2 */, uint32_t *period)
{
    const uint32_t lo = LTP_MIN_PERIOD, hi = LTP_MAX_PERIOD;
    uint32_t cand[20], ncand = 0, i = lo, k;
    double best_peak = 0.0;
    while (i < hi && ncand < 20) {
        uint32_t start, end, j, arg = 0; double peak = 0.0;
        for (start = i; start < hi; start++) { if (r[start - 1] < 0.0 && r[start] > 0.0) { break; } }
        for (end = start + 1; end < hi - 1; end++) { if (r[end] > 0.0 && r[end + 1] < 0.0) { break; } }
        for (j = start; j <= end; j++) {
            if (r[j] > r[j - 1] && r[j] > r[j + 1] && r[j] > peak) { arg = j; peak = r[j]; }
        }
        if (arg) { cand[ncand++] = arg; if (peak > best_peak) { best_peak = peak; } }
        i = end + 1;
    }
    if (!ncand || best_peak < 0.1 * r[0]) { return 0; }
    for (k = 0; k < ncand; k++) { if (r[cand[k]] >= 0.9 * best_peak) { *period = cand[k]; return 1; } }
    return 0;
}

static int ltp_analyse(const double *xd, uint32_t n, uint32_t order, uint32_t *period_out, int32_t *qcoef)
{
    double r[LTP_MAX_PERIOD + 4];
    double A[SO_MAX_LTP_ORDER][SO_MAX_LTP_ORDER], inv_diag[SO_MAX_LTP_ORDER], sol[SO_MAX_LTP_ORDER];
    uint32_t period = 0; int32_t i, j, k; const int32_t dim = (int32_t)order; const double *rhs;
    *period_out = 0;
    memset(r, 0, sizeof(r));
    welch_autocorr(xd, n, r, LTP_MAX_PERIOD + 1);           /* lags 0..262 (beyond the FFT size: stale scratch); 263, 264 are
                                                               never written by any call and read as 0 on a fresh handle */
    if (fabs(r[0]) <= FLT_MIN) { return SO_OK; }
    if (!detect_pitch(r, &period)) { return SO_OK; }
    if (period < order / 2 + 1) { return SO_OK; }
    r[0] *= (1.0 + RIDGE);
    for (i = 0; i < dim; i++) { for (j = 0; j < dim; j++) { A[i][j] = r[(i > j) ? i - j : j - i]; } }
    for (i = 0; i < dim; i++) {
        double s = A[i][i];
        for (k = i - 1; k >= 0; k--) { s -= A[i][k] * A[i][k]; }
        if (s <= 0.0) { return SO_NG; }
        inv_diag[i] = pow(s, -0.5);
        for (j = i + 1; j < dim; j++) {
            s = A[i][j];
            for (k = i - 1; k >= 0; k--) { s -= A[i][k] * A[j][k]; }
            A[j][i] = s * inv_diag[i];
        }
    }
    rhs = &r[period - order / 2];
    for (i = 0; i < dim; i++) {
        double s = rhs[i];
        for (j = i - 1; j >= 0; j--) { s -= A[i][j] * sol[j]; }
        sol[i] = s * inv_diag[i];
    }
    for (i = dim - 1; i >= 0; i--) {
        double s = sol[i];
        for (j = i + 1; j < dim; j++) { s -= A[j][i] * sol[j]; }
        sol[i] = s * inv_diag[i];
    }
    for (i = 0; i < dim; i++) {
        int32_t v = (int32_t)round_half_away(sol[i] * pow(2.0, LTP_COEF_BITS - 1));
        if (v < -(1 << (LTP_COEF_BITS - 1))) { v = -(1 << (LTP_COEF_BITS - 1)); }
        if (v > (1 << (LTP_COEF_BITS - 1)) - 1) { v = (1 << (LTP_COEF_BITS - 1)) - 1; }
        qcoef[dim - 1 - i] = v;                               /* stored reversed (srla_encoder.c:1043-1047) */
    }
    *period_out = period;
    return SO_OK;
}

/* ------------------------------------------------------------------------------------------------
 * residual coder analysis (srla_coder.c:262-347, 349-483)
 * ---------------------------------------------------------------------------------------------- */
static uint32_t rice_param(double mean)                     /* srla_coder.c:262-287 */
{
    const double rho = 1.0 / (1.0 + mean);
    const double v = round_half_away(log2_via_ln(log(0.5127629514437670454896078808815218508243560791015625) / log(1.0 - rho)));
    return (uint32_t)((0 > v) ? 0 : v);
}
static uint32_t recursive_rice_k2(double mean)              /* srla_coder.c:298-324 */
{
    const double g = 0.66794162356 * (1.0 + mean);
    const uint32_t golomb = (uint32_t)((1 > g) ? 1 : g);
    return floor_log2(golomb);
}

typedef struct {
    uint32_t code_type, porder, bits;
    uint32_t max_porder;
    uint8_t  k[1 << LOG2_MAX_PARTS];   /* coding parameter of every partition at `porder` */
} rice_plan;

static void rice_plan_search(const int32_t *res, uint32_t n, rice_plan *plan)
{
    static double mean[LOG2_MAX_PARTS + 1][1 << LOG2_MAX_PARTS];
    uint32_t *u = (uint32_t *)malloc(sizeof(uint32_t) * n);
    uint32_t max_porder = 0, nparts, per, part, i, porder, peak = 0, best_bits = UINT32_MAX, best = 0;
    int32_t lvl;
    while (max_porder < LOG2_MAX_PARTS && (n % (2u << max_porder)) == 0) { max_porder++; }
    nparts = 1u << max_porder; per = n / nparts;
    for (part = 0; part < nparts; part++) {
        double sum = 0.0;
        for (i = 0; i < per; i++) {
            const uint32_t v = zigzag(res[part * per + i]);
            u[part * per + i] = v; sum += v; if (v > peak) { peak = v; }
        }
        mean[max_porder][part] = sum / per;
    }
    for (lvl = (int32_t)max_porder - 1; lvl >= 0; lvl--) {
        for (part = 0; part < (1u << lvl); part++) { mean[lvl][part] = (mean[lvl + 1][2 * part] + mean[lvl + 1][2 * part + 1]) / 2.0; }
    }
    plan->max_porder = max_porder;
    if (peak == 0) { plan->code_type = CODE_ALLZERO; plan->porder = 0; plan->bits = 2; free(u); return; }
    plan->code_type = (mean[0][0] < 2) ? CODE_RICE : CODE_RECURSIVE_RICE;
    for (porder = 0; porder <= max_porder; porder++) {
        const uint32_t len = n >> porder;
        uint32_t bits = LOG2_MAX_PARTS, prev = 0;
        for (part = 0; part < (1u << porder); part++) {
            uint32_t k;
            if (plan->code_type == CODE_RICE) {
                k = rice_param(mean[porder][part]);
                for (i = 0; i < len; i++) { bits += 1 + k + (u[part * len + i] >> k); }
            } else {
                const uint32_t k2 = recursive_rice_k2(mean[porder][part]), k1 = k2 + 1;
                k = k2;
                bits += (k1 + 1) * len;
                for (i = 0; i < len; i++) {
                    const int32_t over = (int32_t)u[part * len + i] - (int32_t)(1u << k1);
                    bits += (uint32_t)((over > 0 ? over : 0) >> k2);
                }
            }
            bits += (part == 0) ? RICE_PARAM_BITS : zigzag((int32_t)k - (int32_t)prev) + 1;
            prev = k;
        }
        if (bits < best_bits) { best_bits = bits; best = porder; }   /* the reference's early exit cannot change this */
    }
    plan->porder = best; plan->bits = best_bits + 2;
    for (part = 0; part < (1u << best); part++) {
        plan->k[part] = (uint8_t)((plan->code_type == CODE_RICE) ? rice_param(mean[best][part]) : recursive_rice_k2(mean[best][part]));
    }
    free(u);
}

uint32_t so_rice_search(const int32_t *residual, uint32_t n, uint32_t *code_type, uint32_t *porder)
{
    rice_plan plan; rice_plan_search(residual, n, &plan);
    if (code_type) { *code_type = plan.code_type; }
    if (porder) { *porder = plan.porder; }
    return plan.bits;
}

/* emission (srla_coder.c:165-190, 486-595) */
static void rice_emit(bitsink *s, const int32_t *res, uint32_t n)
{
    rice_plan plan; uint32_t part, i, len, prev = 0;
    rice_plan_search(res, n, &plan);
    put_bits(s, plan.code_type, 2);
    if (plan.code_type == CODE_ALLZERO) { return; }
    put_bits(s, plan.porder, LOG2_MAX_PARTS);
    len = n >> plan.porder;
    for (part = 0; part < (1u << plan.porder); part++) {
        const uint32_t k = plan.k[part];
        if (part == 0) { put_bits(s, k, RICE_PARAM_BITS); } else { put_zero_run(s, zigzag((int32_t)k - (int32_t)prev)); }
        prev = k;
        for (i = 0; i < len; i++) {
            const uint32_t v = zigzag(res[part * len + i]);
            if (plan.code_type == CODE_RICE) {
                put_zero_run(s, v >> k); put_bits(s, v, k);
            } else {
                const uint32_t k1 = k + 1, pivot = 1u << k1;
                if (v < pivot) { put_bits(s, pivot | v, k1 + 1); }
                else { const uint32_t z = v - pivot; put_zero_run(s, 1 + (z >> k)); put_bits(s, z, k); }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * SVR coefficient refinement (lpc.c:988-1136; call site srla_encoder.c:1087-1101), used when
 * so_set_svr_iterations() > 0: covariance of the un-windowed signal, ridge, Cholesky (lpc.c:573-631, with
 * libm pow(s, -0.5) like the reference), then for six margins up to `iterations` re-weighted least-squares steps
 * on the soft-thresholded residual; the coefficients of the smallest recursive-Rice length estimate win.
 * ---------------------------------------------------------------------------------------------- */
static uint32_t g_svr_iterations = 0;
void so_set_svr_iterations(uint32_t iterations) { g_svr_iterations = iterations; }

static double svr_objective(double mean_abs)                                       /* lpc.c:1020-1030, BITS_PER_SAMPLE 16 */
{
    const double intmean = mean_abs * (1 << 16);
    const double rho = 1.0 / (1.0 + intmean);
    const double l2 = log2_via_ln(log(0.5127629514) / log(1.0 - rho));
    const uint32_t k2 = (uint32_t)((0 > l2) ? 0 : l2);
    const uint32_t k1 = k2 + 1;
    const double k1factor = pow(1.0 - rho, (double)(1 << k1));
    const double k2factor = pow(1.0 - rho, (double)(1 << k2));
    return (1.0 + k1) * (1.0 - k1factor) + (1.0 + k2 + (1.0 / (1.0 - k2factor))) * k1factor;
}

static void svr_refine(const double *data, uint32_t n, double *coef, uint32_t dim, uint32_t iterations)
{
    static const double margins[6] = { 0.0, 1.0 / 4096, 1.0 / 1024, 1.0 / 256, 1.0 / 64, 1.0 / 16 };   /* srla_internal.c:27 */
    double *cov = (double *)calloc((size_t)dim * dim, sizeof(double));
    double *resid = (double *)malloc(sizeof(double) * n);
    double inv_diag[SO_MAX_ORDER], rvec[SO_MAX_ORDER], delta[SO_MAX_ORDER], init[SO_MAX_ORDER], best[SO_MAX_ORDER];
    double min_obj = FLT_MAX;
    uint32_t i, j, smpl, m, itr;
    int k, singular = 0;
#define COV(a, b) cov[(size_t)(a) * dim + (b)]
    for (smpl = 0; smpl < n - dim; smpl++) {
        const double *pd = &data[smpl];
        for (i = 0; i < dim; i++) { const double sv = pd[i]; for (j = i; j < dim; j++) { COV(i, j) += sv * pd[j]; } }
    }
    for (i = 0; i < dim; i++) { for (j = i + 1; j < dim; j++) { COV(j, i) = COV(i, j); } }
    for (i = 0; i < dim; i++) { COV(i, i) *= (1.0 + RIDGE); }
    for (i = 0; i < dim && !singular; i++) {                                        /* lpc.c:573-602 */
        double sum = COV(i, i);
        for (k = (int)i - 1; k >= 0; k--) { sum -= COV(i, k) * COV(i, k); }
        if (sum <= 0.0) { singular = 1; break; }
        inv_diag[i] = pow(sum, -0.5);
        for (j = i + 1; j < dim; j++) {
            sum = COV(i, j);
            for (k = (int)i - 1; k >= 0; k--) { sum -= COV(i, k) * COV(j, k); }
            COV(j, i) = sum * inv_diag[i];
        }
    }
    if (singular) { for (i = 0; i < dim; i++) { coef[i] = 0.0; } free(cov); free(resid); return; }
    memcpy(init, coef, sizeof(double) * dim);
    memcpy(best, coef, sizeof(double) * dim);
    for (m = 0; m < 6; m++) {
        const double margin = margins[m];
        double prev_obj = FLT_MAX;
        memcpy(coef, init, sizeof(double) * dim);
        for (itr = 0; itr < iterations; itr++) {
            double mabse = 0.0, obj;
            memcpy(resid, data, sizeof(double) * n);
            for (i = 0; i < dim; i++) { rvec[i] = 0.0; }
            for (smpl = dim; smpl < n; smpl++) {
                double r;
                for (i = 0; i < dim; i++) { resid[smpl] += coef[i] * data[smpl - i - 1]; }
                r = resid[smpl];
                mabse += (r > 0) ? r : -r;
                { const double mag = ((r > 0) ? r : -r) - margin; resid[smpl] = ((r > 0) - (r < 0)) * ((mag > 0.0) ? mag : 0.0); }
                for (i = 0; i < dim; i++) { rvec[i] += resid[smpl] * data[smpl - i - 1]; }
            }
            obj = svr_objective(mabse / n);
            for (i = 0; i < dim; i++) {                                             /* lpc.c:605-631 */
                double sum = rvec[i];
                for (k = (int)i - 1; k >= 0; k--) { sum -= COV(i, k) * delta[k]; }
                delta[i] = sum * inv_diag[i];
            }
            for (k = (int)dim - 1; k >= 0; k--) {
                double sum = delta[k];
                for (j = (uint32_t)k + 1; j < dim; j++) { sum -= COV(j, k) * delta[j]; }
                delta[k] = sum * inv_diag[k];
            }
            if (obj < min_obj) { memcpy(best, coef, sizeof(double) * dim); min_obj = obj; }
            if ((prev_obj < obj) || (fabs(prev_obj - obj) < 1e-8)) { break; }
            for (i = 0; i < dim; i++) { coef[i] += delta[i]; }
            prev_obj = obj;
        }
    }
    memcpy(coef, best, sizeof(double) * dim);
#undef COV
    free(cov); free(resid);
}

/* ------------------------------------------------------------------------------------------------
 * one candidate channel (srla_encoder.c:966-1205)
 * ---------------------------------------------------------------------------------------------- */
int so_analyse_channel(const so_params *p, int32_t *sig, uint32_t n, int32_t *residual, so_channel *out)
{
    static double rows[SO_MAX_ORDER][SO_MAX_ORDER + 3];
    const uint32_t P = PRESET_MAX_ORDER[p->preset], bps = p->bits_per_sample;
    const double unit = pow(2.0, -(int32_t)(bps - 1));
    const huff_table *plain = format_huffman(0), *summed = format_huffman(1);
    double *xd = (double *)malloc(sizeof(double) * n);
    uint32_t i, order = 0, rshift = 0, bits;
    memset(out, 0, sizeof(*out));

    out->pre_prev = sig[0];
    out->pre_coef = preemphasis_coef(sig, n);
    preemphasis_apply(sig, n, out->pre_coef);

    if (p->ltp_order > 0) {
        int rc;
        for (i = 0; i < n; i++) { xd[i] = sig[i] * unit; }
        rc = ltp_analyse(xd, n, p->ltp_order, &out->ltp_period, out->ltp_coef);
        if (rc != SO_OK) { free(xd); return rc; }
        if (out->ltp_period > 0) {
            ltp_residual(sig, n, out->ltp_coef, p->ltp_order, out->ltp_period, residual);
            memcpy(sig, residual, sizeof(int32_t) * n);
        }
    }

    if (P > 0) {
        for (i = 0; i < n; i++) { xd[i] = sig[i] * unit; }
        welch_autocorr(xd, n, out->autocorr, P + 1);
        out->autocorr[0] *= (1.0 + RIDGE);
        levinson(out->autocorr, P, rows, out->error_vars);
        { const double g = welch_energy_gain(n); for (i = 0; i <= P; i++) { out->error_vars[i] *= g; } }
        order = choose_order(out->error_vars, P, n, bps);    /* every preset > 0 uses the estimation tactic */
    }
    if (order > 0) {
        int32_t q[SO_MAX_ORDER];
        memcpy(out->lpc_double, &rows[order - 1][1], sizeof(double) * order);
        if (g_svr_iterations > 0) { svr_refine(xd, n, out->lpc_double, order, g_svr_iterations); }
        quantise_lpc(out->lpc_double, order, q, &rshift);
        for (i = 0; i < order; i++) { out->coef[i] = q[order - 1 - i]; }   /* FIR order (srla_encoder.c:1104-1108) */
        fir_residual(sig, n, out->coef, order, rshift, residual);
    } else {
        memcpy(residual, sig, sizeof(int32_t) * n);
        rshift = 0;
    }
    out->order = order; out->rshift = rshift;

    out->residual_bits = so_rice_search(residual, n, &out->code_type, &out->porder);
    bits = out->residual_bits;
    bits += bps + 1 + (PRE_SHIFT + 1);
    bits += ORDER_BITS + RSHIFT_BITS + 1;
    if (order > 0) {                                           /* srla_encoder.c:1141-1174 */
        uint32_t plain_bits = 0, sum_bits, use_sum = 1;
        for (i = 0; i < order; i++) { plain_bits += plain->len[zigzag(out->coef[i])]; }
        sum_bits = plain->len[zigzag(out->coef[0])];
        for (i = 1; i < order; i++) {
            const uint32_t sym = zigzag(out->coef[i] + out->coef[i - 1]);
            if (sym >= 256) { use_sum = 0; break; }
            sum_bits += summed->len[sym];
            if (sum_bits >= plain_bits) { use_sum = 0; break; }
        }
        out->use_sum = use_sum;
        bits += use_sum ? sum_bits : plain_bits;
    }
    bits += 1;
    if (out->ltp_period > 0) { bits += 1 + LTP_PERIOD_BITS + p->ltp_order * LTP_COEF_BITS; }
    out->total_bits = bits;
    free(xd);
    return SO_OK;
}

/* ------------------------------------------------------------------------------------------------
 * block level (srla_encoder.c:766-796, 799-858, 1208-1334, 1337-1455, 1477-1643)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t nch, n, method;
    uint32_t payload_bits;   /* what the reference's size estimate returns: channels 0/1 only (srla_encoder.c:1276-1301) */
    uint32_t emitted_bits;   /* what emission really produces: every channel */
    so_channel chan[SO_MAX_CHANNELS];
    int32_t *res[SO_MAX_CHANNELS];
    int32_t *storage;
} block_plan;

static void block_plan_free(block_plan *b) { free(b->storage); b->storage = NULL; }

static uint32_t block_type_of(const so_params *p, const int32_t *const *pcm, uint32_t n)
{
    uint32_t c, i;
    if (n <= PRESET_MAX_ORDER[p->preset]) { return BLOCK_RAW; }
    for (c = 0; c < p->num_channels; c++) { for (i = 0; i < n; i++) { if (pcm[c][i]) { return BLOCK_COMPRESS; } } }
    return BLOCK_SILENT;
}

/* analyse all candidates, choose the stereo method (first minimum over LR, MS, LS, SR) */
static int block_plan_make(const so_params *p, const int32_t *const *pcm, uint32_t n, block_plan *b)
{
    const uint32_t nch = p->num_channels;
    const uint32_t ncand = (nch >= 2) ? nch + 2 : nch;   /* [M, S,] ch0, ch1, ... */
    so_channel *cand = (so_channel *)malloc(sizeof(so_channel) * ncand);
    int32_t *sig = (int32_t *)malloc(sizeof(int32_t) * n);
    uint32_t c, i, first_ch = (nch >= 2) ? 2 : 0;
    int rc = SO_OK;
    b->nch = nch; b->n = n; b->method = 0;
    b->storage = (int32_t *)malloc(sizeof(int32_t) * n * ncand);
    for (c = 0; c < ncand && rc == SO_OK; c++) {
        if (nch >= 2 && c < 2) {
            for (i = 0; i < n; i++) {
                const int32_t l = asr(pcm[0][i], p->offset_lshift), r = asr(pcm[1][i], p->offset_lshift);
                const int32_t side = wrap_sub(r, l);
                sig[i] = (c == 1) ? side : wrap_add(l, asr(side, 1));     /* srla_utility.c:91-103 */
            }
        } else {
            for (i = 0; i < n; i++) { sig[i] = asr(pcm[c - first_ch][i], p->offset_lshift); }
        }
        rc = so_analyse_channel(p, sig, n, b->storage + (size_t)c * n, &cand[c]);
    }
    if (rc == SO_OK) {
        uint32_t map[SO_MAX_CHANNELS];
        for (c = 0; c < nch; c++) { map[c] = first_ch + c; }
        b->payload_bits = 0;
        if (nch >= 2) {
            const uint32_t M = cand[0].total_bits, S = cand[1].total_bits, L = cand[2].total_bits, R = cand[3].total_bits;
            const uint32_t cost[4] = { L + R, M + S, L + S, S + R };
            uint32_t m, best = 0;
            for (m = 1; m < 4; m++) { if (cost[best] > cost[m]) { best = m; } }
            b->method = best;
            if (best == 1) { map[0] = 0; map[1] = 1; } else if (best == 2) { map[1] = 1; } else if (best == 3) { map[0] = 1; }
            /* NB only channels 0/1 enter the returned bit count (srla_encoder.c:1276-1301) */
            b->payload_bits = cost[best];
        } else {
            b->payload_bits = cand[0].total_bits;
        }
        b->payload_bits += 2;
        b->payload_bits = (b->payload_bits + 7u) & ~7u;
        b->emitted_bits = 2;
        for (c = 0; c < nch; c++) {
            b->chan[c] = cand[map[c]]; b->res[c] = b->storage + (size_t)map[c] * n;
            b->emitted_bits += b->chan[c].total_bits;
        }
        b->emitted_bits = (b->emitted_bits + 7u) & ~7u;
    } else {
        block_plan_free(b);
    }
    free(sig); free(cand);
    return rc;
}

static uint32_t raw_bytes(const so_params *p, uint32_t n) { return (p->bits_per_sample * n * p->num_channels) / 8; }

int so_block_size(const so_params *p, const int32_t *const *pcm, uint32_t n, uint32_t *size)
{
    uint32_t type;
    if (!p || !pcm || !n || !size) { return SO_INVALID_ARGUMENT; }
    if (n > p->max_block) { return SO_INSUFFICIENT_BUFFER; }
    type = block_type_of(p, pcm, n);
    if (type == BLOCK_COMPRESS) {
        block_plan b; int rc = block_plan_make(p, pcm, n, &b);
        if (rc != SO_OK) { return rc; }
        block_plan_free(&b);
        if (b.payload_bits >= p->bits_per_sample * n * p->num_channels) { type = BLOCK_RAW; }
        else { *size = BLOCK_HEADER_BYTES + b.payload_bits / 8; return SO_OK; }
    }
    *size = BLOCK_HEADER_BYTES + ((type == BLOCK_RAW) ? raw_bytes(p, n) : 0);
    return SO_OK;
}

static uint32_t emit_compressed(const so_params *p, const block_plan *b, uint8_t *dst)
{
    const huff_table *plain = format_huffman(0), *summed = format_huffman(1);
    bitsink s; uint32_t c, i;
    s.mem = dst; s.bit = 0;
    put_bits(&s, b->method, 2);
    for (c = 0; c < b->nch; c++) {
        put_bits(&s, zigzag(b->chan[c].pre_prev), p->bits_per_sample + 1);
        put_bits(&s, zigzag(b->chan[c].pre_coef), PRE_SHIFT + 1);
    }
    for (c = 0; c < b->nch; c++) {
        const so_channel *ch = &b->chan[c];
        put_bits(&s, ch->order, ORDER_BITS);
        put_bits(&s, ch->rshift, RSHIFT_BITS);
        put_bits(&s, ch->use_sum, 1);
        for (i = 0; i < ch->order; i++) {
            if (i == 0 || !ch->use_sum) { const uint32_t sym = zigzag(ch->coef[i]); put_bits(&s, plain->code[sym], plain->len[sym]); }
            else { const uint32_t sym = zigzag(ch->coef[i] + ch->coef[i - 1]); put_bits(&s, summed->code[sym], summed->len[sym]); }
        }
    }
    for (c = 0; c < b->nch; c++) {
        const so_channel *ch = &b->chan[c];
        put_bits(&s, ch->ltp_period != 0, 1);
        if (ch->ltp_period) {
            put_bits(&s, (p->ltp_order - 1) / 2, 1);
            put_bits(&s, ch->ltp_period - LTP_MIN_PERIOD, LTP_PERIOD_BITS);
            for (i = 0; i < p->ltp_order; i++) { put_bits(&s, zigzag(ch->ltp_coef[i]), LTP_COEF_BITS); }
        }
    }
    for (c = 0; c < b->nch; c++) { rice_emit(&s, b->res[c], b->n); }
    return sink_bytes(&s);
}

int so_encode_block(const so_params *p, const int32_t *const *pcm, uint32_t n, uint8_t *out, uint32_t cap, uint32_t *size)
{
    uint32_t type, payload = 0, c, i;
    if (!p || !pcm || !n || !out || !cap || !size) { return SO_INVALID_ARGUMENT; }
    if (n > p->max_block) { return SO_INSUFFICIENT_BUFFER; }
    type = block_type_of(p, pcm, n);
    if (type == BLOCK_COMPRESS) {
        block_plan b; int rc = block_plan_make(p, pcm, n, &b);
        if (rc != SO_OK) { return rc; }
        /* the encoder compares the bytes it really wrote (srla_encoder.c:1608) */
        if (b.emitted_bits >= p->bits_per_sample * n * p->num_channels) { type = BLOCK_RAW; }
        else {
            if (cap < BLOCK_HEADER_BYTES + b.emitted_bits / 8) { block_plan_free(&b); return SO_INSUFFICIENT_BUFFER; }
            memset(out, 0, BLOCK_HEADER_BYTES + b.emitted_bits / 8);
            payload = emit_compressed(p, &b, out + BLOCK_HEADER_BYTES);
            if (payload * 8 != b.emitted_bits) { block_plan_free(&b); return SO_NG; }   /* self-check */
        }
        block_plan_free(&b);
    }
    if (type == BLOCK_RAW) {
        const int bytes = (int)(p->bits_per_sample / 8);
        uint8_t *w = out + BLOCK_HEADER_BYTES;
        if (cap < BLOCK_HEADER_BYTES + raw_bytes(p, n)) { return SO_INSUFFICIENT_BUFFER; }
        for (i = 0; i < n; i++) { for (c = 0; c < p->num_channels; c++) { store_be(w, zigzag(pcm[c][i]), bytes); w += bytes;   /* zig-zag mapped, srla_encoder.c:826-849 */ } }
        payload = (uint32_t)(w - (out + BLOCK_HEADER_BYTES));
    }
    if (cap < BLOCK_HEADER_BYTES) { return SO_INSUFFICIENT_BUFFER; }
    store_be(out + 0, 0xFFFFu, 2);
    store_be(out + 2, payload + 5, 4);
    out[8] = (uint8_t)type;
    store_be(out + 9, n, 2);
    store_be(out + 6, so_fletcher16(out + 8, payload + 3), 2);
    *size = BLOCK_HEADER_BYTES + payload;
    return SO_OK;
}

/* ------------------------------------------------------------------------------------------------
 * variable block division: exact sizes of every candidate segment + dense shortest path
 * (srla_encoder.c:249-307, 310-424)
 * ---------------------------------------------------------------------------------------------- */
int so_search_partition(const so_params *p, const int32_t *const *pcm, uint32_t n, uint32_t *num_parts, uint32_t *parts)
{
    const uint32_t unit = p->min_block;
    const uint32_t nodes = (n + unit - 1) / unit + 1;
    const double BIG = (double)(1UL << 24);
    double *edge = (double *)malloc(sizeof(double) * nodes * nodes);
    double *dist = (double *)malloc(sizeof(double) * nodes);
    uint32_t *from = (uint32_t *)malloc(sizeof(uint32_t) * nodes);
    uint8_t *done = (uint8_t *)calloc(nodes, 1);
    uint32_t i, j, c, cur = 0, count = 0, node;
    int rc = SO_OK;
    for (i = 0; i < nodes * nodes; i++) { edge[i] = BIG; }
    for (i = 0; i < nodes && rc == SO_OK; i++) {
        for (j = i + 1; j < nodes && rc == SO_OK; j++) {
            const int32_t *seg[SO_MAX_CHANNELS];
            uint32_t len = (j - i) * unit, bytes;
            if (len > p->max_block) { continue; }
            if (len > n - i * unit) { len = n - i * unit; }
            for (c = 0; c < p->num_channels; c++) { seg[c] = pcm[c] + i * unit; }
            rc = so_block_size(p, seg, len, &bytes);
            edge[i * nodes + j] = bytes;
        }
    }
    if (rc == SO_OK) {
        for (i = 0; i < nodes; i++) { dist[i] = BIG; from[i] = ~0u; }
        dist[0] = 0.0;
        for (;;) {
            double low = BIG;
            for (i = 0; i < nodes; i++) { if (!done[i] && low > dist[i]) { low = dist[i]; cur = i; } }
            if (cur == nodes - 1) { break; }
            for (i = 0; i < nodes; i++) {
                if (dist[i] > edge[cur * nodes + i] + dist[cur]) { dist[i] = edge[cur * nodes + i] + dist[cur]; from[i] = cur; }
            }
            done[cur] = 1;
        }
        for (node = nodes - 1; node != 0; node = from[node]) { count++; }
        node = nodes - 1;
        for (i = 0; i < count; i++) {
            uint32_t len = (node - from[node]) * unit;
            if (len > n - from[node] * unit) { len = n - from[node] * unit; }
            parts[count - 1 - i] = len;
            node = from[node];
        }
        *num_parts = count;
    }
    free(edge); free(dist); free(from); free(done);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * whole stream (srla_encoder.c:1701-1788)
 * ---------------------------------------------------------------------------------------------- */
int so_encode_whole(const so_params *p_in, const int32_t *const *pcm, uint32_t num_samples, uint8_t *out, uint32_t cap, uint32_t *size)
{
    so_params p; uint32_t done = 0, at = FILE_HEADER_BYTES, c; int rc;
    if (!p_in || !pcm || !out || !size) { return SO_INVALID_ARGUMENT; }
    p = *p_in;
    so_reset_state();                                           /* a fresh handle (the CLI creates one per file) */
    p.offset_lshift = so_offset_lshift(pcm, p.num_channels, num_samples);
    if ((rc = so_encode_header(&p, num_samples, out, cap)) != SO_OK) { return rc; }
    while (done < num_samples) {
        const int32_t *at_pcm[SO_MAX_CHANNELS];
        const uint32_t step = (p.min_block == p.max_block) ? p.max_block : p.lookahead;
        const uint32_t todo = (step < num_samples - done) ? step : num_samples - done;
        for (c = 0; c < p.num_channels; c++) { at_pcm[c] = pcm[c] + done; }
        if (p.min_block == p.max_block) {
            uint32_t wrote;
            if ((rc = so_encode_block(&p, at_pcm, todo, out + at, cap - at, &wrote)) != SO_OK) { return rc; }
            at += wrote;
        } else {
            uint32_t *parts = (uint32_t *)malloc(sizeof(uint32_t) * (todo / p.min_block + 2));
            uint32_t nparts = 0, k, off = 0;
            rc = so_search_partition(&p, at_pcm, todo, &nparts, parts);
            for (k = 0; k < nparts && rc == SO_OK; k++) {
                const int32_t *part_pcm[SO_MAX_CHANNELS]; uint32_t wrote;
                for (c = 0; c < p.num_channels; c++) { part_pcm[c] = at_pcm[c] + off; }
                rc = so_encode_block(&p, part_pcm, parts[k], out + at, cap - at, &wrote);
                if (rc == SO_OK) { at += wrote; off += parts[k]; }
            }
            free(parts);
            if (rc != SO_OK) { return rc; }
        }
        done += todo;
    }
    *size = at;
    return SO_OK;
}

int so_encode_whole_flat(const so_params *p, const int32_t *pcm, uint32_t num_samples, uint8_t *out, uint32_t cap, uint32_t *size)
{
    const int32_t *rows[SO_MAX_CHANNELS]; uint32_t c;
    if (!p || p->num_channels > SO_MAX_CHANNELS) { return SO_INVALID_ARGUMENT; }
    for (c = 0; c < p->num_channels; c++) { rows[c] = pcm + (size_t)c * num_samples; }
    return so_encode_whole(p, rows, num_samples, out, cap, size);
}

/* ================================================================================================
 * Decoder restatement (srla_decoder.c:63-799, srla_coder.c:596-690, srla_lpc_synthesize.c:238-327,
 * srla_utility.c:106-174, 361-378): checks the GPU decoder when oracle/_ref is not at hand.
 * Decoding by code matching instead of a tree walk: the static Huffman codes are prefix free, so the
 * first (code, length) that matches the next bits is the symbol the reference's tree walk ends at.
 * ============================================================================================== */
typedef struct { const uint8_t *mem; uint64_t bit, limit; int over; } bitsrc;

static uint32_t get_bits(bitsrc *s, uint32_t nbits)
{
    uint32_t v = 0, i;
    for (i = 0; i < nbits; i++) {
        uint32_t b = 0;
        if (s->bit < s->limit) { b = (s->mem[s->bit >> 3] >> (7u - (uint32_t)(s->bit & 7u))) & 1u; } else { s->over = 1; }
        s->bit++;
        v = (v << 1) | b;
    }
    return v;
}
static uint32_t get_zero_run(bitsrc *s) { uint32_t run = 0; while (!s->over && get_bits(s, 1) == 0u) { run++; } return run; }
static int32_t unzigzag(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }            /* srla_utility.h:33 */

static uint32_t huff_get(bitsrc *s, const huff_table *t)
{
    uint32_t code = 0, len, sym;
    for (len = 1; len <= 32 && !s->over; len++) {
        code = (code << 1) | get_bits(s, 1);
        for (sym = 0; sym < 256; sym++) { if (t->len[sym] == len && t->code[sym] == code) { return sym; } }
    }
    s->over = 1;
    return 0;
}

static void residual_decode(bitsrc *s, int32_t *x, uint32_t n)                                   /* srla_coder.c:648-690 */
{
    const uint32_t type = get_bits(s, 2);
    uint32_t porder, per, part, k = 0, i;
    if (type == 2u) { memset(x, 0, sizeof(int32_t) * n); return; }
    if (type > 2u) { s->over = 1; return; }
    porder = get_bits(s, 10);
    if (porder > 10u) { s->over = 1; return; }
    per = n >> porder;
    for (part = 0; part < (1u << porder) && !s->over; part++) {
        if (part == 0) { k = get_bits(s, 5); } else { k = (uint32_t)((int32_t)k + unzigzag(get_zero_run(s))); }
        if (k > 31u) { s->over = 1; return; }
        for (i = 0; i < per && !s->over; i++) {
            const uint32_t quot = get_zero_run(s);
            uint32_t u;
            if (type == 0u) { u = (quot << k) + get_bits(s, k); }
            else { u = get_bits(s, k + (quot ? 0u : 1u)); u |= (quot + (quot ? 1u : 0u)) << k; }
            x[part * per + i] = unzigzag(u);
        }
    }
}

/* one block at data (size bytes available) -> pcm[ch][0..n); returns an SO_* code; *used, *n_out like the reference */
static int decode_block(const so_params *p, int check, const uint8_t *data, uint64_t size, int32_t *const *pcm, uint32_t cap_samples,
                        uint32_t *used, uint32_t *n_out)
{
    uint32_t bsize, n, type, ch, i;
    const uint32_t nch = p->num_channels, bps = p->bits_per_sample;
    if (size < 11u) { return SO_INSUFFICIENT_DATA; }
    if (data[0] != 0xFF || data[1] != 0xFF) { return SO_INVALID_FORMAT; }
    bsize = ((uint32_t)data[2] << 24) | ((uint32_t)data[3] << 16) | ((uint32_t)data[4] << 8) | data[5];
    if ((uint64_t)bsize + 6u > size) { return SO_INSUFFICIENT_DATA; }
    if (bsize < 5u) { return SO_INVALID_FORMAT; }
    if (check && so_fletcher16(data + 8, bsize - 2u) != (uint16_t)(((uint32_t)data[6] << 8) | data[7])) { return SO_DATA_CORRUPTION; }
    type = data[8];
    n = ((uint32_t)data[9] << 8) | data[10];
    if (n > cap_samples) { return SO_INSUFFICIENT_BUFFER; }
    if (type == 2u) {                                                                             /* raw: srla_decoder.c:363-433 */
        const uint32_t sb = bps / 8u;
        const uint8_t *q = data + 11;
        if ((uint64_t)bsize - 5u < ((uint64_t)bps * n * nch) / 8u) { return SO_INSUFFICIENT_DATA; }
        for (i = 0; i < n; i++) { for (ch = 0; ch < nch; ch++) { uint32_t u = 0, b; for (b = 0; b < sb; b++) { u = (u << 8) | *q++; } pcm[ch][i] = unzigzag(u); } }
    } else if (type == 1u) {
        for (ch = 0; ch < nch; ch++) { memset(pcm[ch], 0, sizeof(int32_t) * n); }
    } else if (type == 0u) {
        bitsrc s; uint32_t method;
        int32_t head[SO_MAX_CHANNELS], pre[SO_MAX_CHANNELS], coef[SO_MAX_CHANNELS][SO_MAX_ORDER + 1], ltp_coef[SO_MAX_CHANNELS][4];
        uint32_t order[SO_MAX_CHANNELS], rshift[SO_MAX_CHANNELS], ltp_order[SO_MAX_CHANNELS], ltp_period[SO_MAX_CHANNELS];
        s.mem = data + 11; s.bit = 0; s.limit = 8ull * (bsize - 5u); s.over = 0;
        method = get_bits(&s, 2);
        for (ch = 0; ch < nch; ch++) { head[ch] = unzigzag(get_bits(&s, bps + 1u)); pre[ch] = unzigzag(get_bits(&s, 5)); }
        for (ch = 0; ch < nch; ch++) {
            uint32_t use_sum;
            order[ch] = get_bits(&s, 8); rshift[ch] = get_bits(&s, 4); use_sum = get_bits(&s, 1);
            for (i = 0; i < order[ch]; i++) {
                coef[ch][i] = unzigzag(huff_get(&s, format_huffman(use_sum && i > 0)));
                if (use_sum && i > 0) { coef[ch][i] -= coef[ch][i - 1]; }
            }
        }
        for (ch = 0; ch < nch; ch++) {
            ltp_order[ch] = 0; ltp_period[ch] = 0;
            if (get_bits(&s, 1)) {
                ltp_order[ch] = 2u * get_bits(&s, 1) + 1u; ltp_period[ch] = get_bits(&s, 8) + 8u;
                for (i = 0; i < ltp_order[ch]; i++) { ltp_coef[ch][i] = unzigzag(get_bits(&s, 6)); }
            }
        }
        for (ch = 0; ch < nch && !s.over; ch++) { residual_decode(&s, pcm[ch], n); }
        if (s.over) { return SO_DATA_CORRUPTION; }                      /* the reference has no such check: only reachable with the checksum test off */
        for (ch = 0; ch < nch; ch++) {
            int32_t *x = pcm[ch];
            const uint32_t P = order[ch];
            if (P > 0u) {                                                                         /* srla_lpc_synthesize.c:238-262 */
                const int32_t half = (rshift[ch] > 0u) ? (int32_t)(1u << (rshift[ch] - 1u)) : (int32_t)0x80000000u;
                for (i = 1; i < P && i < n; i++) { x[i] = wrap_add(x[i], x[i - 1]); }
                for (i = 0; n > P && i < n - P; i++) {
                    int32_t predict = half; uint32_t o;
                    for (o = 0; o < P; o++) { predict = wrap_add(predict, wrap_mul(coef[ch][o], x[i + o])); }
                    x[i + P] = wrap_sub(x[i + P], asr(predict, rshift[ch]));
                }
            }
            if (ltp_order[ch] > 0u && ltp_period[ch] > 0u) {                                      /* srla_lpc_synthesize.c:264-327 */
                const uint32_t h = ltp_order[ch] >> 1, T = ltp_period[ch];
                for (i = T + h + 1u; i < n; i++) {
                    int32_t predict = 16; uint32_t o;
                    for (o = 0; o < ltp_order[ch]; o++) { predict = wrap_add(predict, wrap_mul(ltp_coef[ch][o], x[i - T - h + o])); }
                    x[i] = wrap_add(x[i], asr(predict, 5));
                }
            }
            x[0] = wrap_add(x[0], asr(wrap_mul(head[ch], pre[ch]), 4));                           /* srla_utility.c:361-378 */
            for (i = 1; i < n; i++) { x[i] = wrap_add(x[i], asr(wrap_mul(x[i - 1], pre[ch]), 4)); }
        }
        if (nch >= 2u && method != 0u) {                                                          /* srla_utility.c:106-174 */
            for (i = 0; i < n; i++) {
                int32_t a = pcm[0][i], b = pcm[1][i];
                if (method == 1u) { a = wrap_sub(a, asr(b, 1)); b = wrap_add(b, a); }
                else if (method == 2u) { b = wrap_add(b, a); }
                else { a = wrap_sub(b, a); }
                pcm[0][i] = a; pcm[1][i] = b;
            }
        }
        if (p->offset_lshift > 0u) { for (ch = 0; ch < nch; ch++) { for (i = 0; i < n; i++) { pcm[ch][i] = (int32_t)((uint32_t)pcm[ch][i] << p->offset_lshift); } } }
    } else {
        return SO_INVALID_FORMAT;
    }
    *used = bsize + 6u; *n_out = n;
    return SO_OK;
}

int so_decode_header(const uint8_t *data, uint32_t size, so_params *p, uint32_t *num_samples)
{
    uint32_t fmt, codec;
    if (!data || !p || !num_samples) { return SO_INVALID_ARGUMENT; }
    if (size < 30u) { return SO_INSUFFICIENT_DATA; }
    if (data[0] != '1' || data[1] != '2' || data[2] != '4' || data[3] != '9') { return SO_INVALID_FORMAT; }
    fmt = ((uint32_t)data[4] << 24) | ((uint32_t)data[5] << 16) | ((uint32_t)data[6] << 8) | data[7];
    codec = ((uint32_t)data[8] << 24) | ((uint32_t)data[9] << 16) | ((uint32_t)data[10] << 8) | data[11];
    memset(p, 0, sizeof(*p));
    p->num_channels = ((uint32_t)data[12] << 8) | data[13];
    *num_samples = ((uint32_t)data[14] << 24) | ((uint32_t)data[15] << 16) | ((uint32_t)data[16] << 8) | data[17];
    p->sampling_rate = ((uint32_t)data[18] << 24) | ((uint32_t)data[19] << 16) | ((uint32_t)data[20] << 8) | data[21];
    p->bits_per_sample = ((uint32_t)data[22] << 8) | data[23];
    p->offset_lshift = data[24];
    p->max_block = ((uint32_t)data[25] << 24) | ((uint32_t)data[26] << 16) | ((uint32_t)data[27] << 8) | data[28];
    p->min_block = p->max_block; p->lookahead = p->max_block;
    p->preset = data[29];
    /* srla_decoder.c:137-182 (checked by SetHeader there) */
    if (fmt != 10u || codec != 18u || p->num_channels == 0 || *num_samples == 0 || p->sampling_rate == 0 || p->bits_per_sample == 0
        || p->offset_lshift >= 32u || p->max_block == 0 || p->preset >= 7u || p->num_channels > SO_MAX_CHANNELS) { return SO_INVALID_FORMAT; }
    return SO_OK;
}

/* header + every block (srla_decoder.c:740-799); pcm is [channels][capacity_samples] contiguous */
int so_decode_whole_flat(const uint8_t *data, uint32_t size, int check_checksum, int32_t *pcm, uint32_t channels, uint32_t capacity_samples)
{
    so_params p; uint32_t total = 0, progress = 0, ch; uint64_t at = 30;
    int32_t *rows[SO_MAX_CHANNELS];
    int rc = so_decode_header(data, size, &p, &total);
    if (rc != SO_OK) { return rc; }
    if (channels < p.num_channels || capacity_samples < total) { return SO_INSUFFICIENT_BUFFER; }
    while (progress < total && at < size) {
        uint32_t used = 0, n = 0;
        for (ch = 0; ch < p.num_channels; ch++) { rows[ch] = pcm + (size_t)ch * capacity_samples + progress; }
        rc = decode_block(&p, check_checksum, data + at, size - at, rows, capacity_samples - progress, &used, &n);
        if (rc != SO_OK) { return rc; }
        at += used; progress += n;
    }
    return SO_OK;
}
