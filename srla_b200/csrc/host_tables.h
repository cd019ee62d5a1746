/*
 * host_tables.h -- tables the kernels need that must come from the HOST libm / host integer code
 * so that device results are bit-identical to the reference running on the same machine:
 *
 *  - complex-FFT stage twiddles: the reference advances w <- w * (cos t, flag sin t) sequentially
 *    (fft.c:80-87, 107); the sequence is a pure function of the stage size, so it is tabulated
 *    once here with the host's cos/sin and the same multiply order.
 *  - real-FFT split twiddles: the recurrence of fft.c:149-152, 181-183.
 *  - plain-Rice parameter thresholds: k(mean) of srla_coder.c:262-287 uses libm log(); it is
 *    monotone in mean, so the 31 switch-over means are found by bisection against the host libm.
 *  - static Huffman (code, length) tables for quantised LPC coefficients (static_huffman.c:28-132,
 *    frequencies: srla_internal.c:9-25).
 */
#ifndef SRLA_B200_HOST_TABLES_H
#define SRLA_B200_HOST_TABLES_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/srla_format_tables.h"

namespace srla {
namespace host {

constexpr double kPi = 3.14159265358979323846;   /* fft.c:17 */

struct Cx { double re, im; };
static inline Cx cmul(Cx a, Cx b) { Cx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }

/* forward (flag = -1) twiddles for every stage size 4..max_n (powers of two), one structure-of-arrays
 * table per stage size 2^lg at off[lg] (in Cx units): w1[0..n/4), w2[0..n/4), w3[0..n/4) -- consecutive
 * butterflies of a warp read consecutive 16-byte entries. */
static inline void build_complex_twiddles(int max_lg, std::vector<Cx> &tab, std::vector<uint32_t> &off)
{
    off.assign(32, 0);
    tab.clear();
    for (int lg = 2; lg <= max_lg; lg++) {
        const int n = 1 << lg, quarter = n / 4;
        const double theta = 2.0 * kPi / n;
        const int flag = -1;
        Cx step; step.re = std::cos(theta); step.im = flag * std::sin(theta);
        Cx w; w.re = 1.0; w.im = 0.0;
        off[lg] = (uint32_t)tab.size();
        tab.resize(tab.size() + 3 * (size_t)quarter);
        Cx *t = tab.data() + off[lg];
        for (int p = 0; p < quarter; p++) {
            const Cx w2 = cmul(w, w);
            const Cx w3 = cmul(w, w2);
            t[p] = w; t[quarter + p] = w2; t[2 * quarter + p] = w3;
            w = cmul(w, step);
        }
    }
}

/* forward (flag = -1) split twiddles (wr, wi) for i = 1..N/4, every real size N = 4..max_N.
 * The inverse sequence is the exact conjugate (sin is odd, the recurrence is sign-symmetric);
 * verified against the host libm in build_real_twiddles_check(). */
static inline void real_twiddle_sequence(int n, int flag, std::vector<Cx> &seq)
{
    const double theta = flag * 2.0 * kPi / n;
    const double dsin = std::sin(theta);
    const double dcosm1 = std::cos(theta) - 1.0;
    double wr = 1.0 + dcosm1, wi = dsin;
    seq.clear();
    for (int i = 1; i <= (n >> 2); i++) {
        Cx e; e.re = wr; e.im = wi; seq.push_back(e);
        const double keep = wr;
        wr += keep * dcosm1 - wi * dsin;
        wi += wi * dcosm1 + keep * dsin;
    }
}

static inline bool build_real_twiddles(int max_lg, std::vector<Cx> &tab, std::vector<uint32_t> &off)
{
    off.assign(32, 0);
    tab.clear();
    bool conj_ok = true;
    std::vector<Cx> fwd, inv;
    for (int lg = 2; lg <= max_lg; lg++) {
        real_twiddle_sequence(1 << lg, -1, fwd);
        real_twiddle_sequence(1 << lg, +1, inv);
        for (size_t i = 0; i < fwd.size(); i++) { if (fwd[i].re != inv[i].re || fwd[i].im != -inv[i].im) { conj_ok = false; } }
        off[lg] = (uint32_t)tab.size();
        tab.insert(tab.end(), fwd.begin(), fwd.end());
    }
    return conj_ok;
}

/* plain Rice parameter exactly as the reference evaluates it (srla_coder.c:262-287) */
static inline uint32_t rice_param_libm(double mean)
{
    const double rho = 1.0 / (1.0 + mean);
    const double x = std::log(0.5127629514437670454896078808815218508243560791015625) / std::log(1.0 - rho);
    const double l2 = std::log(x) * 1.4426950408889634;
    const double r = (l2 >= 0.0) ? std::floor(l2 + 0.5) : -std::floor(-l2 + 0.5);
    return (uint32_t)((0 > r) ? 0 : r);
}

/* thr[j] (j = 1..31) = smallest double mean with rice_param_libm(mean) >= j; thr[0] = 0 */
static inline void build_rice_thresholds(double thr[32])
{
    thr[0] = 0.0;
    for (uint32_t j = 1; j < 32; j++) {
        double lo = 0.0, hi = 1.0e12;                       /* k(1e12) > 31 */
        if (rice_param_libm(hi) < j) { thr[j] = std::numeric_limits<double>::infinity(); continue; }
        uint64_t a, b;
        std::memcpy(&a, &lo, 8); std::memcpy(&b, &hi, 8);
        while (b - a > 1) {                                   /* positive doubles order like their bit patterns */
            const uint64_t m = a + (b - a) / 2;
            double md; std::memcpy(&md, &m, 8);
            if (rice_param_libm(md) >= j) { b = m; } else { a = m; }
        }
        std::memcpy(&thr[j], &b, 8);
    }
}

/* static Huffman table: merge the two smallest live nodes under (count, index) order, smaller one
 * on the 0 branch, zero counts bumped to one; codes read root -> leaf. */
struct HuffTable { uint32_t code[256]; uint8_t len[256]; };
/* the same tree for decoding (static_huffman.c:145-162 walks it bit by bit): node >= 256 is internal,
 * child[bit][node - 256] is where `bit` leads; leaves are the symbols */
struct HuffTree { uint16_t child[2][256]; uint16_t root; };

static inline void build_huffman(const uint32_t *counts, uint32_t nsym, HuffTable &t, HuffTree *tree = nullptr)
{
    std::vector<uint32_t> weight(2 * nsym + 1);
    std::vector<uint8_t> live(2 * nsym + 1, 0);
    std::vector<uint32_t> child0(2 * nsym + 1), child1(2 * nsym + 1);
    uint32_t total = nsym;
    std::memset(&t, 0, sizeof(t));
    for (uint32_t i = 0; i < nsym; i++) { weight[i] = counts[i] ? counts[i] : 1u; live[i] = 1; }
    int root = 0;
    for (;;) {
        int a = -1, b = -1;
        for (uint32_t i = 0; i < total; i++) {
            if (!live[i]) { continue; }
            if (a < 0 || weight[i] < weight[a]) { b = a; a = (int)i; }
            else if (b < 0 || weight[i] < weight[b]) { b = (int)i; }
        }
        if (b < 0) { root = a; break; }
        weight[total] = weight[a] + weight[b];
        live[total] = 1; live[a] = live[b] = 0;
        child0[total] = (uint32_t)a; child1[total] = (uint32_t)b;
        total++;
    }
    if (tree) {
        std::memset(tree, 0, sizeof(*tree));
        tree->root = (uint16_t)root;
        for (uint32_t i = nsym; i < total; i++) { tree->child[0][i - nsym] = (uint16_t)child0[i]; tree->child[1][i - nsym] = (uint16_t)child1[i]; }
    }
    /* iterative root -> leaf walk */
    struct Item { uint32_t node, code; uint8_t len; };
    std::vector<Item> stack;
    stack.push_back({ (uint32_t)root, 0u, 0 });
    while (!stack.empty()) {
        const Item it = stack.back(); stack.pop_back();
        if (it.node < nsym) { t.code[it.node] = it.code; t.len[it.node] = it.len; continue; }
        stack.push_back({ child0[it.node], (it.code << 1) | 0u, (uint8_t)(it.len + 1) });
        stack.push_back({ child1[it.node], (it.code << 1) | 1u, (uint8_t)(it.len + 1) });
    }
}

static inline void build_format_huffman(HuffTable &plain, HuffTable &summed, HuffTree *plain_tree = nullptr, HuffTree *summed_tree = nullptr)
{
    static const uint32_t f_plain[256] = SRLA_FMT_COEF_SYMBOL_FREQ_INIT;
    static const uint32_t f_summed[256] = SRLA_FMT_SUMMED_COEF_SYMBOL_FREQ_INIT;
    build_huffman(f_plain, 256, plain, plain_tree);
    build_huffman(f_summed, 256, summed, summed_tree);
}

/* Fletcher-16 (srla_utility.c:36-60) */
static inline uint16_t fletcher16(const uint8_t *d, size_t n)
{
    uint32_t lo = 0, hi = 0;
    for (size_t i = 0; i < n; i++) { lo = (lo + d[i]) % 255u; hi = (hi + lo) % 255u; }
    return (uint16_t)((hi << 8) | lo);
}

} // namespace host
} // namespace srla
#endif
