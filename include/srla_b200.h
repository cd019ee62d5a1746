/*
 * srla_b200.h -- C ABI of libsrla_b200.so, the B200-native (sm_100a CUDA) SRLA encode path.
 *
 * Part 1 is the drop-in boundary: the encoder half of the reference's public API with identical
 * names, struct layouts, argument meaning and result codes, so that tools/srla_codec (or any other
 * caller of the reference library) links against this library unchanged.  Each entry cites the
 * reference interface it replaces (paths relative to the reference tree).
 *
 * Part 2 is the batch / device-resident extension (SURVEY.md section 8b "suggested extension"):
 * many files per submission, PCM already in HBM, output left in HBM.
 *
 * There is no CPU fallback: every encode entry point runs CUDA kernels and returns
 * SRLA_APIRESULT_NG (or Create returns NULL) when no usable device is present.
 */
#ifndef SRLA_B200_H_INCLUDED
#define SRLA_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Part 1 -- reference-compatible surface
 * ---------------------------------------------------------------------------------------------- */

/* Format constants: include/srla.h:6-24 */
#define SRLA_FORMAT_VERSION          10
#define SRLA_CODEC_VERSION           18
#define SRLA_HEADER_SIZE             30
#define SRLA_MAX_NUM_CHANNELS        8
#define SRLA_MAX_COEFFICIENT_ORDER   255
#define SRLA_MAX_LTP_ORDER           3
#define SRLA_NUM_PARAMETER_PRESETS   7

/* Result codes: include/srla.h:29-38 (same numeric values) */
typedef enum SRLAApiResultTag {
    SRLA_APIRESULT_OK = 0,
    SRLA_APIRESULT_INVALID_ARGUMENT,
    SRLA_APIRESULT_INVALID_FORMAT,
    SRLA_APIRESULT_INSUFFICIENT_BUFFER,
    SRLA_APIRESULT_INSUFFICIENT_DATA,
    SRLA_APIRESULT_PARAMETER_NOT_SET,
    SRLA_APIRESULT_DETECT_DATA_CORRUPTION,
    SRLA_APIRESULT_NG
} SRLAApiResult;

/* Stream header: include/srla.h:41-51 (field order and types are ABI) */
struct SRLAHeader {
    uint32_t format_version;
    uint32_t codec_version;
    uint16_t num_channels;
    uint32_t num_samples;                 /* per channel */
    uint32_t sampling_rate;
    uint16_t bits_per_sample;
    uint8_t  offset_lshift;               /* common trailing zero bits removed before coding */
    uint32_t max_num_samples_per_block;
    uint8_t  preset;
};

/* Per-stream parameters: include/srla_encoder.h:8-18 */
struct SRLAEncodeParameter {
    uint16_t num_channels;
    uint16_t bits_per_sample;
    uint32_t sampling_rate;
    uint32_t min_num_samples_per_block;
    uint32_t max_num_samples_per_block;
    uint32_t num_lookahead_samples;
    uint32_t ltp_order;                              /* 0 (off), 1 or 3 */
    uint32_t num_svr_filter_learning_iteration;      /* iterations of the SVR coefficient refinement (0: off) */
    uint8_t  preset;                                 /* 0..6 -> max LPC order 0,8,16,32,64,128,255 */
};

/* Handle capacity: include/srla_encoder.h:21-27 */
struct SRLAEncoderConfig {
    uint32_t max_num_channels;
    uint32_t min_num_samples_per_block;
    uint32_t max_num_samples_per_block;
    uint32_t max_num_lookahead_samples;
    uint32_t max_num_parameters;
};

struct SRLAEncoder;   /* opaque */

/* include/srla_encoder.h:35-36.  Invoked once per top-level step (block, or look-ahead chunk in
 * variable-block mode) in stream order with the reference's arguments; because blocks are encoded
 * as one GPU batch the calls are issued after the batch completes, before EncodeWhole returns. */
typedef void (*SRLAEncoder_EncodeBlockCallback)(
    uint32_t num_samples, uint32_t progress_samples, const uint8_t *encoded_block_data, uint32_t block_data_size);

/* include/srla_encoder.h:43-44 (srla_encoder.c:85-165): 30-byte big-endian stream header. Host only. */
SRLAApiResult SRLAEncoder_EncodeHeader(const struct SRLAHeader *header, uint8_t *data, uint32_t data_size);

/* include/srla_encoder.h:47 (srla_encoder.c:468-546): bytes of caller-provided work memory that
 * Create needs for the HOST side of the handle, -1 for an invalid config (same validity rules).
 * Device memory is owned by the handle and is not part of this figure. */
int32_t SRLAEncoder_CalculateWorkSize(const struct SRLAEncoderConfig *config);

/* include/srla_encoder.h:50 (srla_encoder.c:549-694): (work == NULL && work_size == 0) => the
 * handle allocates for itself and Destroy frees; otherwise the caller owns `work`.
 * NULL on bad arguments, short work area, or when no CUDA device is usable.
 * Capacity limit of this implementation: max_num_samples_per_block <= 65535 (the block header's 16-bit sample count; blocks beyond 16384 samples
 * work in global memory instead of shared memory; SVR refinement is limited to 16384). */
struct SRLAEncoder *SRLAEncoder_Create(const struct SRLAEncoderConfig *config, void *work, int32_t work_size);

/* include/srla_encoder.h:53 (srla_encoder.c:697-707) */
void SRLAEncoder_Destroy(struct SRLAEncoder *encoder);

/* include/srla_encoder.h:56-57 (srla_encoder.c:710-763): same validation and result codes;
 * additionally INVALID_FORMAT when bits_per_sample is not one of 8/16/24 (the widths the format's raw
 * blocks can carry, srla_encoder.c:825-852). */
SRLAApiResult SRLAEncoder_SetEncodeParameter(struct SRLAEncoder *encoder, const struct SRLAEncodeParameter *parameter);

/* include/srla_encoder.h:60-62 (srla_encoder.c:1477-1546): exact encoded size of one block. */
SRLAApiResult SRLAEncoder_ComputeBlockSize(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples, uint32_t *output_size);

/* include/srla_encoder.h:65-68 (srla_encoder.c:1549-1643): one block, host planar int32 in, bytes out. */
SRLAApiResult SRLAEncoder_EncodeBlock(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
    uint8_t *data, uint32_t data_size, uint32_t *output_size);

/* include/srla_encoder.h:71-74 (srla_encoder.c:1646-1698): one look-ahead chunk with optimal division. */
SRLAApiResult SRLAEncoder_EncodeOptimalPartitionedBlock(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
    uint8_t *data, uint32_t data_size, uint32_t *output_size);

/* include/srla_encoder.h:77-81 (srla_encoder.c:1701-1788): header + all blocks of a stream.
 * input = planar host int32 PCM (sign-extended), num_samples per channel.
 * encode_callback: one call per block (per look-ahead chunk with variable blocks) in stream order, like
 * srla_encoder.c:1780-1782; long inputs report while later blocks are still being encoded. */
SRLAApiResult SRLAEncoder_EncodeWhole(
    struct SRLAEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
    uint8_t *data, uint32_t data_size, uint32_t *output_size, SRLAEncoder_EncodeBlockCallback encode_callback);

/* ---- decoder: include/srla_decoder.h:8-56, same names, layouts and result codes ---- */

/* include/srla_decoder.h:8-12 */
struct SRLADecoderConfig {
    uint32_t max_num_channels;
    uint32_t max_num_parameters;
    uint8_t  check_checksum;              /* 1: verify every block's Fletcher-16, anything else: do not */
};

struct SRLADecoder;   /* opaque */

/* include/srla_decoder.h:22-23 (srla_decoder.c:63-134): 30-byte header -> struct. Host only. */
SRLAApiResult SRLADecoder_DecodeHeader(const uint8_t *data, uint32_t data_size, struct SRLAHeader *header);

/* include/srla_decoder.h:26 (srla_decoder.c:185-218): host bytes Create needs; -1 for an invalid config. */
int32_t SRLADecoder_CalculateWorkSize(const struct SRLADecoderConfig *config);

/* include/srla_decoder.h:29 (srla_decoder.c:221-312); NULL also when no CUDA device is usable. */
struct SRLADecoder *SRLADecoder_Create(const struct SRLADecoderConfig *config, void *work, int32_t work_size);

/* include/srla_decoder.h:32 (srla_decoder.c:315-322) */
void SRLADecoder_Destroy(struct SRLADecoder *decoder);

/* include/srla_decoder.h:35-36 (srla_decoder.c:325-360); additionally INVALID_FORMAT for bit depths other than 8/16/24. */
SRLAApiResult SRLADecoder_SetHeader(struct SRLADecoder *decoder, const struct SRLAHeader *header);

/* include/srla_decoder.h:39-43 (srla_decoder.c:633-737): one block, host bytes in, planar host int32 out. */
SRLAApiResult SRLADecoder_DecodeBlock(
    struct SRLADecoder *decoder, const uint8_t *data, uint32_t data_size,
    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples,
    uint32_t *decode_size, uint32_t *num_decode_samples);

/* include/srla_decoder.h:46-49 (srla_decoder.c:740-799): header + every block; all blocks of the stream are
 * decoded by one kernel launch (one CTA per block). */
SRLAApiResult SRLADecoder_DecodeWhole(
    struct SRLADecoder *decoder, const uint8_t *data, uint32_t data_size,
    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples);

/* device time (CUDA events) of the decode kernel of the handle's most recent call, in ms */
float SRLAB200_DecoderKernelMs(const struct SRLADecoder *decoder);

/* ------------------------------------------------------------------------------------------------
 * Part 2 -- batch / device-resident extension (not in the reference)
 * ---------------------------------------------------------------------------------------------- */

/* One stream ("file") of a batch.  All streams of a batch share the handle's current encode
 * parameters (channels, bit depth, rate, block sizes, preset, LTP order). */
struct SRLAB200Stream {
    const void *pcm;          /* planar PCM: channel c starts at pcm + c * channel_stride samples   */
    uint64_t channel_stride;  /* in samples                                                         */
    uint32_t num_samples;     /* per channel                                                        */
    uint32_t sample_bytes;    /* 2 = int16_t samples (bits_per_sample <= 16), 4 = int32_t samples   */
};

/* Encode `num_streams` streams whose PCM already lives in device memory; the concatenated .srl
 * streams are written to device memory `d_out` (capacity out_capacity bytes).
 * stream_offsets[num_streams + 1] (host) receives the byte range of each stream inside d_out.
 * All work is enqueued on the handle's stream and completed before return.
 * Returns INSUFFICIENT_BUFFER if out_capacity is too small (nothing useful is written then). */
SRLAApiResult SRLAB200_EncodeStreamsDevice(
    struct SRLAEncoder *encoder, const struct SRLAB200Stream *streams, uint32_t num_streams,
    uint8_t *d_out, uint64_t out_capacity, uint64_t *stream_offsets);

/* Same, but PCM and output are HOST buffers (ideally pinned): PCM is copied to the device, results
 * are copied back.  sample_bytes/pcm describe host memory here. */
SRLAApiResult SRLAB200_EncodeStreamsHost(
    struct SRLAEncoder *encoder, const struct SRLAB200Stream *streams, uint32_t num_streams,
    uint8_t *out, uint64_t out_capacity, uint64_t *stream_offsets);

/* WAV ingest.  One stream whose PCM is still in the shape of a WAV `data` chunk: frames of
 * num_channels little-endian samples of bits_per_sample / 8 bytes each (8-bit: unsigned with offset 128; 16- and
 * 24-bit: two's complement), exactly what the reference's reader consumes one sample at a time before it calls
 * SRLAEncoder_EncodeWhole (libs/wav/src/wav.c:543-553, :841-866; tools/srla_codec/srla_codec.c:103-134). */
struct SRLAB200Frames {
    const void *frames;       /* HOST memory (ideally pinned, see SRLAB200_AllocPinned): the data chunk's payload */
    uint32_t num_samples;     /* frames, i.e. samples per channel                                                 */
};

/* Encode `num_streams` such streams under the handle's current parameters: the payloads are copied to the
 * device as they are, de-interleaved and widened by a CUDA kernel, and encoded like SRLAB200_EncodeStreamsHost
 * does; the outputs are byte-identical to what the reference CLI writes for the same WAV files. */
SRLAApiResult SRLAB200_EncodeInterleavedHost(
    struct SRLAEncoder *encoder, const struct SRLAB200Frames *items, uint32_t num_streams,
    uint8_t *out, uint64_t out_capacity, uint64_t *stream_offsets);

/* Page-locked host memory for the buffers of the *Host entry points (asynchronous copies at full PCIe rate);
 * NULL when the allocation fails. */
void *SRLAB200_AllocPinned(size_t bytes);
void SRLAB200_FreePinned(void *p);

/* Upper bound of the encoded size of a stream of num_samples samples per channel under the
 * handle's current parameters (header + every block stored raw). */
uint64_t SRLAB200_MaxEncodedSize(const struct SRLAEncoder *encoder, uint32_t num_samples);

/* Statistics of the most recent encode call on this handle. */
struct SRLAB200Stats {
    uint64_t num_blocks;          /* blocks emitted                                           */
    uint64_t num_analysed;        /* candidate segments analysed (> num_blocks with -V > 0)   */
    uint64_t kernel_launches;     /* CUDA kernels launched by this library                    */
    uint64_t bytes_in;            /* algorithmic input bytes  (channels x samples x width)    */
    uint64_t bytes_out;           /* encoded bytes incl. headers                              */
    float    ms_analyse;          /* device time of the three analysis kernels (CUDA events)  */
    float    ms_emit;             /* device time of decide + scan + emit kernels              */
    float    ms_total_device;     /* first kernel start -> last kernel end                    */
    uint32_t order_histogram[256];/* chosen LPC order of every emitted channel                */
    uint32_t method_histogram[4]; /* stereo method of every emitted COMPRESS block            */
    uint32_t type_histogram[3];   /* block types: compress, silent, raw                       */
    float    ms_front;            /* front_kernel (mid/side, pre-emphasis, LTP, FFT autocorrelation) */
    float    ms_lpc;              /* lpc_kernel (Levinson-Durbin, order choice, quantisation)  */
    float    ms_residual;         /* residual_kernel (FIR residual, Rice search)               */
};
SRLAApiResult SRLAB200_GetStats(const struct SRLAEncoder *encoder, struct SRLAB200Stats *stats);

/* Device selection for handles created afterwards by this thread (default: current CUDA device). */
SRLAApiResult SRLAB200_SetDevice(int device_ordinal);

/* Multi-GPU hosts (files shard across devices without any exchange): how many CUDA devices there are, and the
 * PCI address ("dddd:bb:dd.f") of one of them (-1: the current device) -- what a host front end needs to place each
 * device's reader / feeder threads and page-locked buffers on the CPUs next to it (/sys/bus/pci/devices/<address>/local_cpulist). */
int SRLAB200_GetDeviceCount(void);
SRLAApiResult SRLAB200_GetDevicePciBusId(int device_ordinal, char *buffer, int buffer_size);

/* Run the handle's kernels on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL
 * restores the handle's own stream).  Lets a host framework time the kernels with its own events. */
SRLAApiResult SRLAB200_SetStream(struct SRLAEncoder *encoder, void *cuda_stream);

/* Library identification string ("srla_b200 <ver> sm_100a ..."). */
const char *SRLAB200_Version(void);

/* Host-only test hook: the int32 -> int16 narrowing the EncodeWhole feeder uses for <= 16-bit sources.
 * Returns nonzero when a sample lies outside int16 (dst is then unspecified). */
uint32_t SRLAB200_TestNarrow(const int32_t *src, int16_t *dst, uint32_t count);

/* ---- stage-level entry point used by the parity tests (host buffers in/out, runs on the GPU) ---- */

/* Analysis of one candidate channel (srla_encoder.c:966-1205) under the handle's current bit depth,
 * preset and LTP order: sig[n] in, residual[n] out, and the decisions in `result`. */
struct SRLAB200ChannelResult {
    int32_t  pre_coef, pre_prev;
    uint32_t order, rshift, use_sum;
    int32_t  coef[SRLA_MAX_COEFFICIENT_ORDER];
    uint32_t ltp_period;
    int32_t  ltp_coef[SRLA_MAX_LTP_ORDER];
    uint32_t code_type, porder, residual_bits, total_bits;
    double   autocorr[SRLA_MAX_COEFFICIENT_ORDER + 1];
    double   error_vars[SRLA_MAX_COEFFICIENT_ORDER + 1];
    double   lpc_double[SRLA_MAX_COEFFICIENT_ORDER];
};
SRLAApiResult SRLAB200_TestAnalyseChannel(
    struct SRLAEncoder *encoder, const int32_t *sig, uint32_t n, int32_t *residual,
    struct SRLAB200ChannelResult *result);

#ifdef __cplusplus
}
#endif
#endif /* SRLA_B200_H_INCLUDED */
