/*
 * srla_b200_batch -- many-file front end of the B200 SRLA encode path.
 *
 * The reference has one production caller of the encoder, `srla -e in.wav out.srl`
 * (tools/srla_codec/srla_codec.c:75-158): it parses the WAV one sample at a time through a bit buffer
 * (libs/wav/src/wav.c:543-553), encodes that one file, writes it.  This tool takes the same encode
 * options (same letters, defaults and range checks, srla_codec.c:38-63, :311-403) and any number of WAV
 * files: headers are parsed on the host with the reference reader's rules, the data chunks are read by a
 * thread team straight into page-locked memory, files of equal format are submitted together through
 * SRLAB200_EncodeInterleavedHost (de-interleaving, widening and the whole encode run on the GPU) and the
 * .srl files -- byte-identical to the reference CLI's -- are written by a second team; reading, encoding
 * and writing of successive submissions overlap.
 *
 * Host-only C++: it binds nothing but the C ABI of include/srla_b200.h.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/srla_b200.h"

namespace {

struct WavInfo {
    std::string path, out_path;
    uint32_t channels = 0, rate = 0, bits = 0, frames = 0;
    uint64_t data_at = 0, data_bytes = 0, file_bytes = 0;
    bool ok = false;
    uint64_t encoded = 0;
};

uint32_t le(const unsigned char *p, int n) { uint32_t v = 0; for (int i = 0; i < n; i++) { v |= (uint32_t)p[i] << (8 * i); } return v; }

/* Header walk with the rules of the reference reader (wav.c:136-281): "RIFF" <size> "WAVE", then the "fmt "
 * chunk FIRST, of size 16 (format tag 1) or 40 (tag 0xFFFE, extension size 22), then chunks are skipped by their
 * stated size (no padding byte, like the reference's seek) until "data"; frames = data bytes / frame bytes. */
bool parse_wav(WavInfo &w, std::string &why)
{
    FILE *fp = std::fopen(w.path.c_str(), "rb");
    if (!fp) { why = "cannot open"; return false; }
    struct stat sb;
    if (fstat(fileno(fp), &sb) != 0) { std::fclose(fp); why = "cannot stat"; return false; }
    w.file_bytes = (uint64_t)sb.st_size;
    unsigned char h[64];
    auto need = [&](size_t n) { return std::fread(h, 1, n, fp) == n; };
    bool good = false;
    do {
        if (!need(12) || std::memcmp(h, "RIFF", 4) != 0 || std::memcmp(h + 8, "WAVE", 4) != 0) { why = "not a RIFF/WAVE file"; break; }
        if (!need(8) || std::memcmp(h, "fmt ", 4) != 0) { why = "fmt chunk does not follow the WAVE signature"; break; }
        const uint32_t fmt_size = le(h + 4, 4);
        if (fmt_size != 16 && fmt_size != 40) { why = "unsupported fmt chunk size"; break; }
        if (!need(fmt_size)) { why = "truncated fmt chunk"; break; }
        const uint32_t tag = le(h, 2);
        if ((fmt_size == 16 && tag != 1) || (fmt_size == 40 && tag != 0xFFFE)) { why = "not linear PCM"; break; }
        w.channels = le(h + 2, 2); w.rate = le(h + 4, 4); w.bits = le(h + 14, 2);
        if (fmt_size == 40 && le(h + 16, 2) != 22) { why = "bad WAVEFORMATEXTENSIBLE extension size"; break; }
        for (;;) {
            if (!need(8)) { why = "no data chunk"; break; }
            const uint32_t size = le(h + 4, 4);
            if (std::memcmp(h, "data", 4) == 0) { w.data_at = (uint64_t)std::ftell(fp); w.data_bytes = size; good = true; break; }
            std::fprintf(stderr, "WARNING: skiping chunk:%.4s size:%d \n", (const char *)h, (int32_t)size);
            if (std::fseek(fp, (long)(int32_t)size, SEEK_CUR) != 0) { why = "seek failed"; break; }
        }
    } while (0);
    std::fclose(fp);
    if (!good) { return false; }
    if (w.bits != 8 && w.bits != 16 && w.bits != 24) { why = "unsupported bits per sample"; return false; }   /* the format's raw blocks carry 8/16/24 */
    if (w.channels == 0 || w.channels > SRLA_MAX_NUM_CHANNELS) { why = "unsupported channel count"; return false; }
    const uint32_t frame = (w.bits / 8) * w.channels;
    w.frames = (uint32_t)(w.data_bytes / frame);
    if (w.frames == 0) { why = "empty data chunk"; return false; }
    if (w.data_at + (uint64_t)w.frames * frame > w.file_bytes) { why = "data chunk is longer than the file"; return false; }
    return true;
}

/* run fn(i) for i in [0, count) on `threads` host threads */
template <class F> void parallel_for(size_t count, int threads, F fn)
{
    std::atomic<size_t> next{0};
    std::vector<std::thread> team;
    const int t = (int)std::min<size_t>((size_t)std::max(1, threads), std::max<size_t>(count, 1));
    for (int k = 0; k < t; k++) { team.emplace_back([&] { for (;;) { const size_t i = next.fetch_add(1); if (i >= count) { break; } fn(i); } }); }
    for (std::thread &th : team) { th.join(); }
}

bool read_range(const std::string &path, uint64_t at, unsigned char *dst, uint64_t bytes)
{
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { return false; }
    uint64_t done = 0;
    while (done < bytes) {
        const ssize_t got = pread(fd, dst + done, (size_t)std::min<uint64_t>(bytes - done, 1u << 30), (off_t)(at + done));
        if (got <= 0) { break; }
        done += (uint64_t)got;
    }
    close(fd);
    return done == bytes;
}

bool write_file(const std::string &path, const uint8_t *src, uint64_t bytes)
{
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp) { return false; }
    const bool ok = std::fwrite(src, 1, (size_t)bytes, fp) == bytes;
    return (std::fclose(fp) == 0) && ok;
}

bool parse_u32(const char *prog, const char *what, const char *str, uint32_t *out)
{
    char *e = nullptr;
    const long v = std::strtol(str, &e, 10);
    if (*e != '\0' || e == str) { std::fprintf(stderr, "%s: invalid %s. (irregular character found in %s at %s)\n", prog, what, str, e); return false; }
    if (v < 0 || v > 0x7fffffffL) { std::fprintf(stderr, "%s: invalid %s. (%s is out of range)\n", prog, what, str); return false; }
    *out = (uint32_t)v;
    return true;
}

std::vector<int> parse_cpulist(const std::string &text)
{
    std::vector<int> cpus;
    size_t at = 0;
    while (at < text.size()) {
        size_t end = text.find(',', at);
        if (end == std::string::npos) { end = text.size(); }
        const std::string part = text.substr(at, end - at);
        const size_t dash = part.find('-');
        if (!part.empty() && part[0] >= '0' && part[0] <= '9') {
            const int a = std::atoi(part.c_str()), b = (dash == std::string::npos) ? a : std::atoi(part.c_str() + dash + 1);
            for (int c = a; c <= b; c++) { cpus.push_back(c); }
        }
        at = end + 1;
    }
    return cpus;
}

std::vector<int> device_local_cpus(int dev)
{
    char bdf[32] = { 0 };
    if (SRLAB200_GetDevicePciBusId(dev, bdf, (int)sizeof(bdf)) != SRLA_APIRESULT_OK) { return {}; }
    for (char *q = bdf; *q; q++) { if (*q >= 'A' && *q <= 'F') { *q = (char)(*q - 'A' + 'a'); } }
    const std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/local_cpulist";
    FILE *fp = std::fopen(path.c_str(), "r");
    if (!fp) { return {}; }
    char line[4096] = { 0 };
    const bool got = std::fgets(line, sizeof(line), fp) != nullptr;
    std::fclose(fp);
    return got ? parse_cpulist(line) : std::vector<int>();
}

/* The calling thread (one per GPU) and every thread it starts afterwards -- its reader / writer teams, the library's
 * feeder pool -- stay on the CPUs next to that GPU; the devices that report the same CPU set split it, so the
 * pipelines of different GPUs do not compete for cores and their page-locked buffers are first touched on the GPU's
 * own memory node. */
void pin_thread_next_to_device(int dev, const std::vector<int> &devices, size_t index)
{
    const std::vector<int> mine = device_local_cpus(dev);
    if (mine.empty()) { return; }
    size_t sharers = 0, my_rank = 0;
    for (size_t k = 0; k < devices.size(); k++) {
        if (device_local_cpus(devices[k]) == mine) { if (k == index) { my_rank = sharers; } sharers++; }
    }
    if (sharers == 0) { return; }
    cpu_set_t allowed, set;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) { return; }
    std::vector<int> usable;
    for (int c : mine) { if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) { usable.push_back(c); } }
    if (usable.empty()) { return; }
    const size_t per = std::max<size_t>(1, usable.size() / sharers), lo = std::min(usable.size() - 1, my_rank * per);
    const size_t hi = (my_rank + 1 == sharers) ? usable.size() : std::min(usable.size(), lo + per);
    CPU_ZERO(&set);
    for (size_t k = lo; k < std::max(hi, lo + 1); k++) { CPU_SET(usable[k], &set); }
    (void)sched_setaffinity(0, sizeof(set), &set);
}

void usage(const char *prog)
{
    std::fprintf(stderr,
        "Usage: %s [options] -o OUTPUT_DIR INPUT.wav [INPUT.wav ...]\n"
        "  -m, --mode N                       compress mode 0(fast) .. 6(high compression) (default:4)\n"
        "  -B, --max-block-size N             max number of block samples (default:4096)\n"
        "  -V, --variable-block-divisions N   number of variable block-size divisions (default:1)\n"
        "  -L, --lookahead-sample-factor N    multiply factor for lookahead samples (default:4)\n"
        "  -P, --long-term-prediction N       long term prediction order, odd (default:0, disabled)\n"
        "      --svr-filter-learning-iteration N   iterations of the SVR filter refinement (default:0)\n"
        "  -o, --output-dir DIR               INPUT.wav is written to DIR/INPUT.srl\n"
        "  -j, --threads N                    host threads reading / writing files (default:8)\n"
        "  -g, --device N                     CUDA device ordinal (default: current)\n"
        "      --devices all|N                every CUDA device of the host, each with its own read / encode / write pipeline\n"
        "      --batch-megabytes N            PCM submitted per GPU call (default:32; page-locking costs ~0.7 ms per MB)\n"
        "      --timing                       print the time spent per stage to stderr\n"
        "Every file is encoded exactly as `srla -e` with the same options would encode it.\n", prog);
}

} // namespace

int main(int argc, char **argv)
{
    const char *prog = argv[0];
    uint32_t mode = 4, max_block = 4096, divisions = 1, factor = 4, ltp = 0, svr = 0, threads = 8, batch_mb = 32;
    bool timing = false, all_devices = false;
    int device = -1;
    std::string out_dir;
    std::vector<std::string> inputs;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&](const char *name) -> const char * { if (i + 1 >= argc) { std::fprintf(stderr, "%s: option %s needs an argument. \n", prog, name); std::exit(1); } return argv[++i]; };
        if (a == "-h" || a == "--help") { usage(prog); return 0; }
        else if (a == "-v" || a == "--version") { std::printf("%s\n", SRLAB200_Version()); return 0; }
        else if (a == "-e" || a == "--encode") { /* the only mode */ }
        else if (a == "-m" || a == "--mode") {
            if (!parse_u32(prog, "encode preset number", value("mode"), &mode)) { return 1; }
            if (mode >= SRLA_NUM_PARAMETER_PRESETS) { std::fprintf(stderr, "%s: encode preset number is out of range. \n", prog); return 1; }
        } else if (a == "-L" || a == "--lookahead-sample-factor") {
            if (!parse_u32(prog, "number of lookahead samples", value("lookahead-sample-factor"), &factor)) { return 1; }
            if (factor == 0 || factor >= (1u << 16)) { std::fprintf(stderr, "%s: lookahead factor is out of range. \n", prog); return 1; }
        } else if (a == "-B" || a == "--max-block-size") {
            if (!parse_u32(prog, "number of block samples", value("max-block-size"), &max_block)) { return 1; }
            if (max_block == 0 || max_block >= (1u << 16)) { std::fprintf(stderr, "%s: number of block samples is out of range. \n", prog); return 1; }
        } else if (a == "-V" || a == "--variable-block-divisions") {
            if (!parse_u32(prog, "number of variable block divisions", value("variable-block-divisions"), &divisions)) { return 1; }
        } else if (a == "-P" || a == "--long-term-prediction") {
            if (!parse_u32(prog, "number of long term prediction order", value("long-term-prediction"), &ltp)) { return 1; }
            if (ltp > 0 && (ltp % 2) == 0) { std::fprintf(stderr, "%s: long term prediction order is must be odd. \n", prog); return 1; }
            if (ltp > SRLA_MAX_LTP_ORDER) { std::fprintf(stderr, "%s: long term prediction order is too large. \n", prog); return 1; }
        } else if (a == "--svr-filter-learning-iteration") {
            if (!parse_u32(prog, "number of lookahead samples", value("svr-filter-learning-iteration"), &svr)) { return 1; }     /* (sic: the reference's message, srla_codec.c:392) */
        } else if (a == "-o" || a == "--output-dir") { out_dir = value("output-dir"); }
        else if (a == "-j" || a == "--threads") { if (!parse_u32(prog, "thread count", value("threads"), &threads)) { return 1; } }
        else if (a == "-g" || a == "--device") { uint32_t d = 0; if (!parse_u32(prog, "device ordinal", value("device"), &d)) { return 1; } device = (int)d; }
        else if (a == "--devices") {
            const std::string v = value("devices");
            if (v == "all") { all_devices = true; }
            else { uint32_t d = 0; if (!parse_u32(prog, "device ordinal", v.c_str(), &d)) { return 1; } device = (int)d; }
        }
        else if (a == "--timing") { timing = true; }
        else if (a == "--batch-megabytes") { if (!parse_u32(prog, "batch size", value("batch-megabytes"), &batch_mb)) { return 1; } }
        else if (!a.empty() && a[0] == '-') { std::fprintf(stderr, "%s: unknown option %s\n", prog, a.c_str()); usage(prog); return 1; }
        else { inputs.push_back(a); }
    }
    if (divisions >= 32 || (max_block >> divisions) == 0) { std::fprintf(stderr, "%s: number of variable block divisions is too large. \n", prog); return 1; }
    if (inputs.empty()) { std::fprintf(stderr, "%s: input file must be specified. \n", prog); return 1; }
    if (out_dir.empty()) { std::fprintf(stderr, "%s: output directory must be specified. \n", prog); return 1; }
    threads = std::max(1u, std::min(64u, threads));
    batch_mb = std::max(1u, batch_mb);
    if (!all_devices && device >= 0 && SRLAB200_SetDevice(device) != SRLA_APIRESULT_OK) { std::fprintf(stderr, "%s: no CUDA device %d. \n", prog, device); return 1; }
    mkdir(out_dir.c_str(), 0777);

    /* ---- headers ---- */
    std::vector<WavInfo> files(inputs.size());
    int failures = 0;
    for (size_t i = 0; i < inputs.size(); i++) {
        WavInfo &w = files[i];
        w.path = inputs[i];
        std::string base = w.path.substr(w.path.find_last_of('/') == std::string::npos ? 0 : w.path.find_last_of('/') + 1);
        const size_t dot = base.find_last_of('.');
        if (dot != std::string::npos && dot > 0) { base.resize(dot); }
        w.out_path = out_dir + "/" + base + ".srl";
        std::string why;
        w.ok = parse_wav(w, why);
        if (!w.ok) { std::fprintf(stderr, "Failed to open %s. (%s)\n", w.path.c_str(), why.c_str()); failures++; }
    }
    /* two inputs that map to the same DIR/NAME.srl (a/x.wav and b/x.wav, x.wav and x.wave) would overwrite each other,
     * possibly from two writer threads at once: the later ones are refused */
    {
        std::vector<size_t> by_out(files.size());
        for (size_t i = 0; i < files.size(); i++) { by_out[i] = i; }
        std::stable_sort(by_out.begin(), by_out.end(), [&](size_t a, size_t b) { return files[a].out_path < files[b].out_path; });
        for (size_t k = 1; k < by_out.size(); k++) {
            WavInfo &w = files[by_out[k]];
            if (w.out_path == files[by_out[k - 1]].out_path && w.ok) {
                std::fprintf(stderr, "%s: %s and %s would both be written to %s; the latter is skipped. \n", prog, files[by_out[k - 1]].path.c_str(), w.path.c_str(), w.out_path.c_str());
                w.ok = false; failures++;
            }
        }
    }

    if (failures == (int)files.size()) { return 1; }          /* nothing to encode: do not even start the device */

    /* ---- files of equal (channels, bits, rate) are submitted together ---- */
    const auto t_begin = std::chrono::steady_clock::now();
    auto seconds_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };

    std::vector<size_t> order;
    for (size_t i = 0; i < files.size(); i++) { if (files[i].ok) { order.push_back(i); } }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        const WavInfo &x = files[a], &y = files[b];
        if (x.channels != y.channels) { return x.channels < y.channels; }
        if (x.bits != y.bits) { return x.bits < y.bits; }
        return x.rate < y.rate;
    });
    auto payload_bytes = [](const WavInfo &w) { return (uint64_t)w.frames * w.channels * (w.bits / 8); };
    auto padded = [](uint64_t v) { return (v + 255u) / 256u * 256u; };

    /* submissions: runs of equally formatted files of about --batch-megabytes of PCM (a file is never split: its
     * header and offset shift need all of it) */
    struct Batch { size_t begin, end; uint64_t bytes; };
    std::vector<Batch> batches;
    const uint64_t batch_bytes = (uint64_t)batch_mb << 20;
    for (size_t at = 0; at < order.size();) {
        const WavInfo &first = files[order[at]];
        Batch b{ at, at, 0 };
        while (b.end < order.size()) {
            const WavInfo &w = files[order[b.end]];
            if (w.channels != first.channels || w.bits != first.bits || w.rate != first.rate) { break; }
            if (b.end > b.begin && b.bytes + payload_bytes(w) > batch_bytes) { break; }
            b.bytes += padded(payload_bytes(w));
            b.end++;
        }
        batches.push_back(b);
        at = b.end;
    }

    /* ---- devices: files are independent, so every GPU runs its own reader -> encoder -> writer pipeline on its own
     * handle and takes the next submission nobody has taken yet; nothing is exchanged between them ---- */
    std::vector<int> devices;
    if (all_devices) {
        const int count = SRLAB200_GetDeviceCount();
        for (int d = 0; d < count; d++) { devices.push_back(d); }
        if (devices.empty()) { std::fprintf(stderr, "%s: no CUDA device. \n", prog); return 1; }
    } else {
        devices.push_back(device);                                    /* -1: the current device */
    }
    if (devices.size() > batches.size() && !batches.empty()) { devices.resize(batches.size()); }
    std::atomic<size_t> next_batch{0};
    std::atomic<int> fatal{0}, failures_atomic{0};
    std::mutex totals_m;
    uint64_t total_in = 0, total_out = 0, total_samples = 0;
    size_t succeeded = 0;
    double t_create = 0.0, t_read = 0.0, t_write = 0.0, t_encode = 0.0, t_pin_in = 0.0, t_pin_out = 0.0, t_first = 0.0;
    const int io_threads = std::max(1, (int)threads / 2);

    auto run_device = [&](const int dev, const size_t dev_index) {
        if (dev >= 0 && SRLAB200_SetDevice(dev) != SRLA_APIRESULT_OK) { std::fprintf(stderr, "%s: no CUDA device %d. \n", prog, dev); fatal.store(1); return; }
        if (devices.size() > 1) { pin_thread_next_to_device(dev, devices, dev_index); }      /* its reader / writer / feeder threads inherit the mask */
        const auto t_dev = std::chrono::steady_clock::now();
        struct SRLAEncoderConfig config;
        config.max_num_channels = SRLA_MAX_NUM_CHANNELS;
        config.min_num_samples_per_block = max_block >> divisions;
        config.max_num_samples_per_block = max_block;
        config.max_num_lookahead_samples = factor * max_block;
        config.max_num_parameters = SRLA_MAX_COEFFICIENT_ORDER;
        struct SRLAEncoder *encoder = SRLAEncoder_Create(&config, NULL, 0);
        if (encoder == NULL) { std::fprintf(stderr, "Failed to create encoder handle. \n"); fatal.store(1); return; }
        double d_create = seconds_since(t_dev), d_read = 0.0, d_write = 0.0, d_encode = 0.0, d_pin_in = 0.0, d_pin_out = 0.0, d_first = 0.0;

        /* Three stages on three slots of page-locked memory, so that reading batch k+1, encoding batch k and writing
         * batch k-1 overlap: reader (thread team) -> encoder (this thread, the only one that touches the handle) ->
         * writer (thread team). */
        struct Slot {
            unsigned char *pcm = nullptr; uint64_t pcm_cap = 0;
            uint8_t *out = nullptr; uint64_t out_cap = 0;
            size_t batch = 0; bool read_ok = true, encoded = false;
            std::vector<struct SRLAB200Frames> items; std::vector<uint64_t> offsets;
        };
        constexpr int kSlots = 3;
        Slot slots[kSlots];
        struct Queue {
            std::mutex m; std::condition_variable cv; std::deque<int> q;
            void push(int v) { { std::lock_guard<std::mutex> g(m); q.push_back(v); } cv.notify_one(); }
            int pop() { std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return !q.empty(); }); const int v = q.front(); q.pop_front(); return v; }
        } free_q, ready_q, done_q;                                      /* -1 = no more submissions */
        for (int i = 0; i < kSlots; i++) { free_q.push(i); }
        auto grow = [&](void **p, uint64_t *cap, uint64_t want) -> bool {
            if (*cap >= want) { return true; }
            SRLAB200_FreePinned(*p);
            *cap = want + want / 8;
            *p = SRLAB200_AllocPinned(*cap);
            if (*p == nullptr) { *cap = 0; std::fprintf(stderr, "%s: cannot allocate %llu MB of page-locked memory. \n", prog, (unsigned long long)(want >> 20)); return false; }
            return true;
        };

        std::thread reader([&] {
            for (;;) {
                const size_t bi = next_batch.fetch_add(1);
                if (bi >= batches.size()) { break; }
                const int si = free_q.pop();
                Slot &sl = slots[si];
                const Batch &b = batches[bi];
                const auto t0 = std::chrono::steady_clock::now();
                sl.batch = bi; sl.read_ok = true; sl.encoded = false;
                const bool grown = !fatal.load() && grow((void **)&sl.pcm, &sl.pcm_cap, b.bytes);
                d_pin_in += seconds_since(t0);
                if (!grown) { fatal.store(1); sl.read_ok = false; ready_q.push(si); continue; }
                const size_t count = b.end - b.begin;
                sl.items.assign(count, SRLAB200Frames{});
                sl.offsets.assign(count + 1, 0);
                uint64_t o = 0;
                for (size_t k = 0; k < count; k++) { const WavInfo &w = files[order[b.begin + k]]; sl.items[k].frames = sl.pcm + o; sl.items[k].num_samples = w.frames; o += padded(payload_bytes(w)); }
                std::atomic<int> bad{0};
                parallel_for(count, io_threads, [&](size_t k) {
                    const WavInfo &w = files[order[b.begin + k]];
                    if (!read_range(w.path, w.data_at, (unsigned char *)sl.items[k].frames, payload_bytes(w))) { std::fprintf(stderr, "Failed to open %s. (read error)\n", w.path.c_str()); bad.store(1); }
                });
                if (bad.load()) { sl.read_ok = false; }
                d_read += seconds_since(t0);
                ready_q.push(si);
            }
            ready_q.push(-1);
        });

        std::thread writer([&] {
            for (;;) {
                const int si = done_q.pop();
                if (si < 0) { break; }
                Slot &sl = slots[si];
                const Batch &b = batches[sl.batch];
                const size_t count = b.end - b.begin;
                const auto t0 = std::chrono::steady_clock::now();
                if (!sl.encoded) {
                    for (size_t k = 0; k < count; k++) { files[order[b.begin + k]].ok = false; failures_atomic++; }
                } else {
                    std::vector<int> wrote(count, 0);
                    parallel_for(count, io_threads, [&](size_t k) {
                        WavInfo &w = files[order[b.begin + k]];
                        w.encoded = sl.offsets[k + 1] - sl.offsets[k];
                        wrote[k] = write_file(w.out_path, sl.out + sl.offsets[k], w.encoded) ? 1 : 0;
                    });
                    std::lock_guard<std::mutex> g(totals_m);
                    for (size_t k = 0; k < count; k++) {
                        WavInfo &w = files[order[b.begin + k]];
                        if (!wrote[k]) { std::fprintf(stderr, "File output error! %s \n", w.out_path.c_str()); w.ok = false; failures_atomic++; continue; }
                        std::printf("finished: %s %llu -> %llu (%6.2f %%) \n", w.path.c_str(), (unsigned long long)w.file_bytes, (unsigned long long)w.encoded,
                                    100.0 * (double)w.encoded / (double)w.file_bytes);
                        total_in += w.file_bytes; total_out += w.encoded; total_samples += (uint64_t)w.frames * w.channels; succeeded++;
                    }
                }
                d_write += seconds_since(t0);
                free_q.push(si);
            }
        });

        /* while the reader page-locks and fills the first slot: one silent block through the handle, so that the
         * kernels are loaded and the small device buffers exist before the first real submission arrives */
        if (!batches.empty()) {
            const WavInfo &first = files[order[batches[0].begin]];
            struct SRLAEncodeParameter parameter;
            parameter.num_channels = (uint16_t)first.channels; parameter.bits_per_sample = (uint16_t)first.bits; parameter.sampling_rate = first.rate;
            parameter.min_num_samples_per_block = max_block >> divisions; parameter.max_num_samples_per_block = max_block;
            parameter.num_lookahead_samples = factor * max_block; parameter.num_svr_filter_learning_iteration = svr;
            parameter.ltp_order = ltp; parameter.preset = (uint8_t)mode;
            if (SRLAEncoder_SetEncodeParameter(encoder, &parameter) == SRLA_APIRESULT_OK) {
                const auto tw = std::chrono::steady_clock::now();
                std::vector<unsigned char> quiet((size_t)max_block * first.channels * (first.bits / 8), first.bits == 8 ? 128 : 0);
                quiet[quiet.size() / 2] ^= 1;                                   /* not a SILENT block: the analysis kernels run */
                std::vector<uint8_t> sink((size_t)SRLAB200_MaxEncodedSize(encoder, max_block));
                struct SRLAB200Frames one; one.frames = quiet.data(); one.num_samples = max_block;
                uint64_t ends[2] = { 0, 0 };
                (void)SRLAB200_EncodeInterleavedHost(encoder, &one, 1, sink.data(), sink.size(), ends);
                d_first = seconds_since(tw);
            }
        }
        for (;;) {
            const int si = ready_q.pop();
            if (si < 0) { break; }
            Slot &sl = slots[si];
            const Batch &b = batches[sl.batch];
            const WavInfo &first = files[order[b.begin]];
            const auto t0 = std::chrono::steady_clock::now();
            if (sl.read_ok && !fatal.load()) {
                struct SRLAEncodeParameter parameter;
                parameter.num_channels = (uint16_t)first.channels;
                parameter.bits_per_sample = (uint16_t)first.bits;
                parameter.sampling_rate = first.rate;
                parameter.min_num_samples_per_block = max_block >> divisions;
                parameter.max_num_samples_per_block = max_block;
                parameter.num_lookahead_samples = factor * max_block;
                parameter.num_svr_filter_learning_iteration = svr;
                parameter.ltp_order = ltp;
                parameter.preset = (uint8_t)mode;
                const SRLAApiResult set = SRLAEncoder_SetEncodeParameter(encoder, &parameter);
                if (set != SRLA_APIRESULT_OK) { std::fprintf(stderr, "Failed to set encode parameter: %d \n", (int)set); }
                else {
                    uint64_t cap = 0;
                    for (size_t k = b.begin; k < b.end; k++) { cap += SRLAB200_MaxEncodedSize(encoder, files[order[k]].frames); }
                    const auto tp = std::chrono::steady_clock::now();
                    const bool grown = grow((void **)&sl.out, &sl.out_cap, cap);
                    d_pin_out += seconds_since(tp);
                    if (!grown) { fatal.store(1); }
                    else {
                        const SRLAApiResult rc = SRLAB200_EncodeInterleavedHost(encoder, sl.items.data(), (uint32_t)(b.end - b.begin), sl.out, sl.out_cap, sl.offsets.data());
                        if (rc != SRLA_APIRESULT_OK) { std::fprintf(stderr, "Failed to encode data: %d \n", (int)rc); }
                        else { sl.encoded = true; }
                    }
                }
            }
            d_encode += seconds_since(t0);
            done_q.push(si);
        }
        done_q.push(-1);
        reader.join();
        writer.join();
        /* the page-locked slots and the handle are reclaimed with the process: unpinning half a gigabyte costs more than
         * the encode */
        std::lock_guard<std::mutex> g(totals_m);
        t_create = std::max(t_create, d_create); t_read = std::max(t_read, d_read); t_write = std::max(t_write, d_write);
        t_encode = std::max(t_encode, d_encode); t_pin_in = std::max(t_pin_in, d_pin_in); t_pin_out = std::max(t_pin_out, d_pin_out); t_first = std::max(t_first, d_first);
    };

    if (devices.size() == 1) { run_device(devices[0], 0); }
    else {
        std::vector<std::thread> team;
        for (size_t k = 0; k < devices.size(); k++) { team.emplace_back(run_device, devices[k], k); }
        for (std::thread &t : team) { t.join(); }
    }
    failures += failures_atomic.load();
    const double wall = seconds_since(t_begin);
    /* the page-locked slots and the handle are reclaimed with the process: unpinning half a gigabyte costs more than
     * the encode */
    std::printf("total: %zu files, %llu -> %llu bytes (%6.2f %%), %.1f Msamples/s in the encode calls (host buffers in, bytes out), %.1f Msamples/s with device start-up and file I/O\n",
                succeeded, (unsigned long long)total_in, (unsigned long long)total_out,
                total_in ? 100.0 * (double)total_out / (double)total_in : 0.0,
                t_encode > 0 ? (double)total_samples / t_encode / 1e6 : 0.0, wall > 0 ? (double)total_samples / wall / 1e6 : 0.0);
    if (timing) {
        std::fprintf(stderr, "[timing] %zu device(s), slowest of each: start-up + handle %.3f s | %zu submissions | read %.3f s (page-locking %.3f) | encode %.3f s (page-locking %.3f) + warm-up %.3f s | write %.3f s (stages overlap) | total %.3f s\n",
                     devices.size(), t_create, batches.size(), t_read, t_pin_in, t_encode, t_pin_out, t_first, t_write, wall);
    }
    std::fflush(stdout);
    if (fatal.load()) { return 1; }
    return failures ? 1 : 0;
}
