/*
 * vdec.h — batched C ABI of the B200-native Viterbi decoder.
 *
 * This is the drop-in boundary for the Viterbi hot path of
 * thomas-emig/convolutional-codes.  The reference decodes one frame per
 * `struct decoder` through six functions (AWGN-channel/include/decoder.h:17-26,
 * binary-symmetric-channel/include/decoder.h:17-26); that legacy surface is
 * re-exported unchanged by libvdec_awgn.so / libvdec_bsc.so
 * (include/vdec_legacy_awgn.h, include/vdec_legacy_bsc.h).  The legacy ABI is
 * one-frame-per-call and 8-bit sized, so throughput goes through the batched
 * entry points below: same code description, same input and output formats,
 * `nframes` independent frames per call.
 *
 * Formats (identical to what the reference decoder consumes / produces):
 *   soft input   float32 [nframes][T][M]   M = 2^n demapper distances per trellis step
 *                                          (common/demapper.c:66-70 -> viterbi-decoder.c:224-236)
 *   quantised    uint8   [nframes][T][M]   VDEC_SOFT_U8: integer branch metrics, same order as the soft input
 *   hard input   uint8   [nframes][T]      one n-bit symbol per step in the low bits
 *                                          (binary-symmetric-channel/viterbi-decoder.c:234-243)
 *   output       uint8   [nframes][(T+7)/8] decoded bits MSB-first, ALL T = L+K-1 bits
 *                                          (tail included, as viterbi-decoder.c:81-89 writes them)
 *   path metric  int32   [nframes]         hard decoder only: final minimum metric
 *                                          (binary-symmetric-channel/viterbi-decoder.c:75-96,143)
 *
 * All functions return 0 on success and -1 on failure, as the reference API does;
 * vdec_last_error() gives the reason.  There is no CPU fallback: without a CUDA
 * device every compute entry point fails.
 *
 * CUDA device: a handle belongs to vdec_config.device; every entry point that takes
 * a handle makes that device current (cudaSetDevice) and leaves it current.
 * vdec_sweep, which drives several devices, puts the caller's device back.
 */
#ifndef VDEC_H
#define VDEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDEC_MAX_POLY 8

enum vdec_metric {
    VDEC_SOFT_F32 = 0,   /* AWGN-channel/viterbi-decoder.c: float path metrics, +INF = unreachable */
    VDEC_HARD_U16 = 1,   /* binary-symmetric-channel/viterbi-decoder.c: uint16 metrics clamped at 0xFF00 */
    VDEC_SOFT_U8  = 2    /* quantised soft mode: the soft decoder of AWGN-channel/viterbi-decoder.c
                            fed with small non-negative INTEGER branch metrics (e.g. common/hard-demapper.c's
                            {0,1,2}); input uint8 [nframes][T][M].  On such inputs the reference's float
                            arithmetic is exact, so the integer kernels are bit-identical to it. */
};

enum vdec_parity {
    VDEC_PARITY_REF = 0,     /* trellis labels exactly as the reference's `(0x6996 >> val) & 1`
                                evaluates on x86-64 (viterbi-decoder.c:52-57, encoder.c:95-100) */
    VDEC_PARITY_LINEAR = 1   /* true GF(2) parity: the textbook code for the same polynomials */
};

enum vdec_flags {
    VDEC_FORCE_GENERIC = 1u,    /* never pick a code-specialised kernel (testing) */
    VDEC_PREFER_LANEGROUP = 2u, /* K = 8 .. 12: take the shuffle-exchange lane-group kernel instead of the
                                   shared-memory register-tile kernel (A/B measurements, testing) */
    VDEC_PREFER_32BIT = 4u,     /* hard decisions: take the 32-bit register kernel instead of the packed 16-bit one
                                   (A/B measurements, testing) */
    VDEC_NO_JIT = 8u            /* polynomials outside the built-in registry: do not compile a specialised kernel at
                                   run time (NVRTC), run the table-driven kernels (also: environment VDEC_JIT=0) */
};

/* Mirrors the fields of `struct code_param` the decoders read (common/include/code.h:9-19),
 * widened to 32 bits.  polynomials[] are left-aligned: bit 63 taps the newest input bit
 * (common/codebook.c:14-17). */
struct vdec_config {
    uint32_t symlen_out;               /* n, 1..3 */
    uint32_t constr_len;               /* K, 2..15 */
    uint32_t block_len;                /* L >= 1 */
    uint32_t metric;                   /* enum vdec_metric */
    uint32_t parity;                   /* enum vdec_parity */
    uint32_t flags;                    /* enum vdec_flags */
    int32_t  device;                   /* CUDA device ordinal */
    uint32_t max_frames;               /* frames per kernel launch; 0 = default */
    uint64_t polynomials[VDEC_MAX_POLY];
    uint32_t metric_max;               /* VDEC_SOFT_U8: largest branch metric the caller will pass (0 = 255);
                                          T * metric_max < 16384 selects the packed 16-bit kernels */
    uint32_t reserved;
};

typedef struct vdec vdec_t;

/* lifecycle — replaces decoder_create + decoder_init (viterbi-decoder.c:144-150,168-206) */
int         vdec_create(vdec_t** out, const struct vdec_config* cfg);
void        vdec_destroy(vdec_t* h);
const char* vdec_last_error(void);

/* host-only helpers (no device needed): the label table / kernel family a config would get */
int         vdec_config_labels(const struct vdec_config* cfg, uint8_t* labels /* [2^(K-1)][2] */);
const char* vdec_config_kernel_name(const struct vdec_config* cfg);

/* Polynomials outside the built-in registry get a kernel specialised at run time (NVRTC; 2-10 s on the first
 * vdec_create for a code, cached per process).  This compiles that kernel WITHOUT loading or running it (no
 * device needed): returns the cubin size, 0 when the config is served without run-time compilation, -1 on failure
 * (vdec_last_error() has the compiler log). */
long        vdec_jit_check(const struct vdec_config* cfg);
/* same for the raw-I/Q variant of the register kernel (K <= 7 soft codes outside the registry); 0 = not applicable */
long        vdec_jit_check_iq(const struct vdec_config* cfg);

/* geometry */
uint32_t    vdec_steps(const vdec_t* h);             /* T = L + K - 1 */
uint32_t    vdec_states(const vdec_t* h);            /* S = 2^(K-1) */
size_t      vdec_in_bytes_per_frame(const vdec_t* h);
size_t      vdec_out_bytes_per_frame(const vdec_t* h);
const char* vdec_kernel_name(const vdec_t* h);       /* which CUDA kernel family serves this code */
/* trellis labels: labels[2*s+i] = expected symbol for (state s, input i) (viterbi-decoder.c:45-63) */
int         vdec_labels(const vdec_t* h, uint8_t* labels);

/* decode `nframes` frames; buffers in device memory, asynchronous on `cuda_stream`
 * (a cudaStream_t, NULL = legacy default stream).  path_metric_dev may be NULL.
 * Soft (float) inputs must be 16-byte aligned (checked: -1 otherwise).
 * ONE CALL IN FLIGHT PER HANDLE: the survivor scratch and staging buffers belong to the handle, so the device entry
 * points of one handle must all be issued on the same stream (or be separated by a synchronisation); concurrent
 * streams need one handle each, as concurrent frames need one `struct decoder` each in the reference. */
int vdec_decode_device(vdec_t* h, const void* in_dev, uint64_t nframes, uint8_t* out_dev,
                       int32_t* path_metric_dev, void* cuda_stream);
/* same with HOST buffers: H2D, kernels, D2H, synchronous (pipelined in chunks).
 * This is the call a reference maintainer binds; the legacy shims use it. */
int vdec_decode_host(vdec_t* h, const void* in_host, uint64_t nframes, uint8_t* out_host,
                     int32_t* path_metric_host);
/* Raw-I/Q entry: replaces demapper_input + decoder_input together.  Input is what
 * the channel delivers to the demapper (AWGN-channel/main.c:104): float32 [nframes][T][2] (I, Q) per
 * trellis step; the branch metrics are formed on the GPU bit-exactly as common/demapper.c:66-70 does
 * (constellation of common/constellations.c for symlen_out = 1..3) and decoded as above.  Halves the
 * bytes that cross PCIe per decoded bit (8.2 instead of 16.2 at K=7, n=2).  Soft metric only. */
int vdec_decode_iq_device(vdec_t* h, const float* iq_dev, uint64_t nframes, uint8_t* out_dev, void* cuda_stream);
int vdec_decode_iq_host(vdec_t* h, const float* iq_host, uint64_t nframes, uint8_t* out_host);
/* Which kernels serve the raw-I/Q entry of this handle: the register kernels (K <= 7) form the branch metrics inside
 * the decoder kernel's input feeder (one launch, 8 B of HBM per trellis step); other families run the demapper kernel
 * into a staging buffer first.  Precondition of the fused form: I/Q finite, |I|, |Q| < 1e19. */
const char* vdec_iq_kernel_name(vdec_t* h);
/* Sliding-window decoding of ONE long or unterminated symbol stream (north_star piece 4).
 * The reference keeps the survivors of the whole frame (viterbi-decoder.c:183) and cannot do this.  The stream
 * of `nsteps` trellis steps (same per-step format as the batched input) is decoded as overlapping windows of
 * `block_steps` payload steps with `depth` extra steps on either side (warm-up from an unknown start state /
 * traceback convergence), all windows in parallel, and the payload bits are stitched into
 * out_dev[(nsteps+7)/8] (MSB-first, bit t = step t).  Equal to full-length decoding whenever the survivor
 * paths merge within `depth` steps (depth >= ~10*K in practice); the final state is the lowest-index argmin
 * as in viterbi-decoder.c:72-79.  Needs a non-catastrophic code (codebook entries 0 and 3 are catastrophic: their
 * polynomials share the factor 1+D and survivor paths need not merge).  block_steps, depth: multiples of 8,
 * depth <= block_steps, block_steps + 2*depth <= 8192. */
enum vdec_stream_flags { VDEC_STREAM_START_KNOWN = 1u /* the stream begins in state 0 (a tail-terminated frame) */ };
int vdec_stream_decode_device(vdec_t* h, const void* in_dev, uint64_t nsteps, uint32_t block_steps, uint32_t depth,
                              uint32_t flags, uint8_t* out_dev, void* cuda_stream);
int vdec_sync(vdec_t* h);
/* Bounds check of the kernels' writes without external tools: a handle created with the environment variable
 * VDEC_GUARD=1 puts 4 KiB guard zones around every buffer the library allocates for itself (survivor scratch, metric
 * staging, sliding-window staging); this synchronises and returns the number of guard bytes that were overwritten
 * (0 = clean, -1 = not a guarded handle).  Callers guard their own input / output buffers the same way
 * (tests/test_gpu_parity.py::test_no_out_of_bounds_writes). */
long vdec_debug_check(vdec_t* h);
/* number of this library's kernels launched through `h` so far */
uint64_t vdec_launch_count(const vdec_t* h);

/* ---- the stages either side of the decoder ("next") ---------------------- */

/* error counting of the drivers' output callbacks (AWGN-channel/main.c:129-141):
 * counters_dev[0] += bits compared (nframes*L), [1] += bit errors, [2] += frames with >= 1 error.
 * decoded rows are vdec_out_bytes_per_frame() apart, payload rows (L+7)/8. */
int vdec_count_errors_device(vdec_t* h, const uint8_t* decoded_dev, const uint8_t* payload_dev,
                             uint64_t nframes, uint64_t* counters_dev, void* cuda_stream);

/* soft demapper, bit-exact to common/demapper.c:43-45,66-70: iq [nsym][2] -> metrics [nsym][2^n] */
int vdec_demap_device(uint32_t symlen_out, const float* iq_dev, uint64_t nsym, float* metrics_dev,
                      void* cuda_stream);

/* hard demapper, bit-exact to common/hard-demapper.c:62-102: each point is snapped to the nearest constellation point
 * (strict '<' on the squared distances: lowest index wins) and the soft demapper's expression is evaluated at the
 * snapped point: iq [nsym][2] -> metrics [nsym][2^n] (the `*_h` curves of results/awgn_channel.m:16-28) */
int vdec_hard_demap_device(uint32_t symlen_out, const float* iq_dev, uint64_t nsym, float* metrics_dev,
                           void* cuda_stream);

/* GPU-resident synthetic link for frames [first_frame, first_frame+nframes): counter-based
 * random payload -> encoder (encoder.c:84-119) -> mapper (mapper.c:54-71) -> AWGN of standard
 * deviation `sigma` per real dimension (main.c:100-102) -> demapper.  Random numbers depend only
 * on (seed, global frame index, step), so results do not depend on how frames are sharded.
 * payload_dev [nframes][(L+7)/8]; metrics_dev [nframes][T][M] and iq_dev [nframes][T][2]: either may be NULL
 * (not both); I/Q only is the cheap form, to be decoded through vdec_decode_iq_device. */
int vdec_synth_awgn_device(vdec_t* h, uint64_t nframes, uint64_t first_frame, uint64_t seed, float sigma,
                           uint8_t* payload_dev, float* metrics_dev, float* iq_dev, void* cuda_stream);
/* binary symmetric channel: each coded bit flipped with probability crossover_ppm/1e6
 * (binary-symmetric-channel/main.c:62-68).  symbols_dev [nframes][T]. */
int vdec_synth_bsc_device(vdec_t* h, uint64_t nframes, uint64_t first_frame, uint64_t seed,
                          uint32_t crossover_ppm, uint8_t* payload_dev, uint8_t* symbols_dev,
                          void* cuda_stream);

/* ---- the drivers' operating-point loop on all GPUs of the node, from C --------------------
 * Replaces the Eb/N0 loop of AWGN-channel/main.c:193-234 and the crossover-rate loop of
 * binary-symmetric-channel/main.c:147-176: per point, batches of `frames_per_batch` synthetic frames are split into
 * contiguous ranges of the global frame index over `ndev` devices (single process, one stream per device), each
 * device runs source -> channel -> [demapper ->] decoder -> error count on its range, and the three counters are
 * summed with one ncclAllReduce(uint64, sum) per batch (NCCL is opened with dlopen; not needed for ndev = 1).
 * A point ends after one batch, or — with min_frame_errors > 0 — once that many errored frames were seen or
 * max_frames were simulated.  Results do not depend on ndev (data is keyed by the global frame index).
 *   points[i]      Eb/N0 in dB (AWGN channels, sigma as main.c:153-161) or the crossover rate in ppm (BSC, main.c:103-109)
 *   devices        CUDA ordinals, NULL = 0 .. ndev-1; ndev = 0: every device present
 *   counters_out   [npoints][3] = bits compared, bit errors, frames in error (main.c:129-141)
 * cfg->metric must be VDEC_SOFT_F32 for the AWGN channels and VDEC_HARD_U16 for the BSC; cfg->device is ignored.
 * Calls from several host threads are serialised (the NCCL communicators are created once per device list and kept). */
enum vdec_channel {
    VDEC_CHANNEL_AWGN = 0,        /* soft demapper (common/demapper.c), fused into the decoder where the kernel family allows */
    VDEC_CHANNEL_AWGN_HARD = 1,   /* hard demapper (common/hard-demapper.c): the `*_h` curves */
    VDEC_CHANNEL_BSC = 2          /* bit flips on the encoder output, hard-decision decoder */
};
struct vdec_sweep_config {
    uint64_t seed;
    uint64_t frames_per_batch;    /* frames per point and batch, all devices together */
    uint64_t max_frames;          /* with min_frame_errors: give up on a point after this many frames (0 = no limit: an error-free point never ends) */
    uint64_t min_frame_errors;    /* 0 = one batch per point */
    uint32_t chunk_frames;        /* frames per launch and device, 0 = 303104 (four full waves of the register kernels) */
    uint32_t channel;             /* enum vdec_channel */
    const char* resume_path;      /* optional: counters are appended here after every batch and picked up on restart */
};
int vdec_sweep(const struct vdec_config* cfg, const struct vdec_sweep_config* sweep, const double* points,
               uint32_t npoints, const int* devices, int ndev, uint64_t* counters_out);
int vdec_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VDEC_H */
