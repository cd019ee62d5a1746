"""Host-side tables of the product (srla_b200/csrc/host_tables.h), checked on the CPU through a tiny C++ harness:
the static Huffman trees the GPU decoder walks must invert the code tables the GPU encoder emits with, the
Fletcher-16 helper must pass the reference's known answers (test/srla_internal/main.cpp:27-29,55-56), the plain-Rice
thresholds must be monotone and reproduce the reference's libm formula at and next to every switch-over point."""
import os
import subprocess

from helpers import ROOT

HARNESS = r'''
#include <cstdio>
#include <cstring>
#include <cmath>
#include "srla_b200/csrc/host_tables.h"
using namespace srla::host;
int main() {
    HuffTable plain, summed; HuffTree t0, t1;
    build_format_huffman(plain, summed, &t0, &t1);
    const HuffTable *tabs[2] = { &plain, &summed }; const HuffTree *trees[2] = { &t0, &t1 };
    for (int k = 0; k < 2; k++) {
        double kraft = 0.0;
        for (int s = 0; s < 256; s++) {
            const uint32_t code = tabs[k]->code[s]; const int len = tabs[k]->len[s];
            if (len < 1 || len > 32) { std::printf("bad length %d\n", len); return 1; }
            kraft += std::ldexp(1.0, -len);
            uint32_t node = trees[k]->root;
            for (int b = len - 1; b >= 0; b--) {
                if (node < 256) { std::printf("leaf reached early: table %d symbol %d\n", k, s); return 1; }
                node = trees[k]->child[(code >> b) & 1u][node - 256];
            }
            if (node != (uint32_t)s) { std::printf("tree walk of table %d symbol %d ends at %u\n", k, s, node); return 1; }
        }
        if (kraft != 1.0) { std::printf("Kraft sum %.17g\n", kraft); return 1; }
    }
    /* Fletcher-16 known answers */
    struct { const char *text; unsigned expect; } kat[] = { { "abcde", 0xC8F0 }, { "abcdef", 0x2057 }, { "abcdefgh", 0x0627 } };
    for (auto &c : kat) { if (fletcher16((const uint8_t *)c.text, std::strlen(c.text)) != c.expect) { std::printf("fletcher %s\n", c.text); return 1; } }
    /* plain-Rice thresholds */
    double thr[32]; build_rice_thresholds(thr);
    for (int j = 1; j < 32; j++) {
        if (!(thr[j] > thr[j - 1])) { std::printf("thresholds not increasing at %d\n", j); return 1; }
        if (std::isinf(thr[j])) { continue; }
        if (rice_param_libm(thr[j]) < (uint32_t)j) { std::printf("threshold %d too low\n", j); return 1; }
        if (rice_param_libm(std::nextafter(thr[j], 0.0)) >= (uint32_t)j) { std::printf("threshold %d not the smallest\n", j); return 1; }
    }
    /* FFT twiddle tables: the inverse real-split sequence is the conjugate of the forward one on this libm */
    std::vector<Cx> tab; std::vector<uint32_t> off;
    if (!build_real_twiddles(14, tab, off)) { std::printf("real twiddles are not conjugate-symmetric\n"); return 1; }
    build_complex_twiddles(12, tab, off);
    if (off[12] + 3u * 1024u != tab.size()) { std::printf("complex twiddle layout\n"); return 1; }
    std::printf("ok\n");
    return 0;
}
'''


def test_host_tables_harness(tmp_path):
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "harness"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I", ROOT, "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr
