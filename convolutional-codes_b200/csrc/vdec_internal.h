// vdec_internal.h — shared between the translation units of libvdec.so (not installed).
//
// The kernel headers (vdec_regs.cuh, vdec_tile.cuh, ...) are also compiled at run time by NVRTC for codes outside the
// built-in registries (vdec_jit.cu); under __CUDACC_RTC__ only the device-side declarations below are visible.
#pragma once
#ifdef __CUDACC_RTC__
#include "vdec.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/vdec.h"
#endif
#include "vdec_labels.h"

namespace vd {

#ifndef __CUDACC_RTC__
void set_error(const char* fmt, ...);

#define VDEC_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ::vd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

#endif  // !__CUDACC_RTC__

// Geometry of one decode launch, passed by value to the kernels.
struct LaunchGeom {
    uint32_t K, n, L, S, T, M;
    uint32_t out_bytes;      // (T+7)/8
    uint32_t nw;             // survivor words (32 decisions each) per trellis step
    uint32_t nframes;        // frames in this launch
    uint32_t stride_steps;   // trellis steps between the starts of consecutive frames in the input
                             // (T for packed frames; smaller for overlapping windows of one stream)
    uint32_t init_all_from;  // frames with index >= this start with ALL states at metric 0 (unknown start
                             // state, sliding-window blocks); 0xFFFFFFFF = every frame starts in state 0
};

#ifndef __CUDACC_RTC__
// ---- code-specialised register kernels (vdec_regs.cu) -------------------------------------
// One thread per frame, all S <= 64 path metrics in registers, labels are compile-time
// constants.  An entry serves a request when its label table equals the requested code's.
struct RegsEntry {
    const char* name;
    uint32_t K, n, metric;
    uint8_t labels[128];                       // [S][2]
    int (*launch)(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric,
                  uint32_t* surv, cudaStream_t st);
    // soft entries only: the same kernel fed with raw I/Q [frame][step][2], branch metrics formed in the feeder
    int (*launch_iq)(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric,
                     uint32_t* surv, cudaStream_t st);
    size_t (*surv_bytes)(const LaunchGeom& g);
};
const RegsEntry* regs_find(uint32_t K, uint32_t n, uint32_t metric, const uint8_t* labels);
int regs_count();
const RegsEntry* regs_at(int i);

// ---- packed 16-bit kernels, two frames per thread (vdec_regs2.cu) ---------------------------------
struct Regs2Entry;
const Regs2Entry* regs2_find(uint32_t K, uint32_t n, uint32_t mode, const uint8_t* labels);
const char* regs2_name(const Regs2Entry* e);
int regs2_run(const Regs2Entry* e, const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric,
              uint32_t* scratch, cudaStream_t st);
size_t regs2_bytes(const Regs2Entry* e, const LaunchGeom& g);

// ---- lane-group register kernels for K = 8 .. 12 (vdec_regsw.cu) -------------------------------------
struct WideEntry;
const WideEntry* regsw_find(uint32_t K, uint32_t n, uint32_t metric, const uint8_t* labels);
const char* regsw_name(const WideEntry* e);
int regsw_run(const WideEntry* e, const LaunchGeom& g, const void* in, uint8_t* out, uint32_t* scratch, cudaStream_t st);
size_t regsw_bytes(const WideEntry* e, const LaunchGeom& g);

// ---- register-tile kernels for K = 9 .. 15 (vdec_tile.cu): 2^(K-7) threads per frame, shared-memory transposes ----
struct TileEntry;
const TileEntry* tile_find(uint32_t K, uint32_t n, uint32_t metric, const uint8_t* labels);
const char* tile_name(const TileEntry* e);
int tile_run(const TileEntry* e, const LaunchGeom& g, const void* in, uint8_t* out, uint32_t* scratch, cudaStream_t st);
size_t tile_bytes(const TileEntry* e, const LaunchGeom& g);

// ---- packed 16-bit register-tile kernels for K = 8 .. 15 (vdec_tile2.cu): hard decisions, quantised soft ----
struct Tile2Entry;
const Tile2Entry* tile2_find(uint32_t K, uint32_t n, uint32_t mode, const uint8_t* labels);
const char* tile2_name(const Tile2Entry* e);
int tile2_run(const Tile2Entry* e, const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch,
              cudaStream_t st);
size_t tile2_bytes(const Tile2Entry* e, const LaunchGeom& g);

// ---- run-time compiled instantiations of the four families above (vdec_jit.cu, NVRTC) -----------------------
struct JitEntry;
enum { kJitRegs = 0, kJitRegs2 = 1, kJitTile = 2, kJitTile2 = 3, kJitRegsIq = 4 /* kJitRegs soft with the raw-I/Q feeder */ };
bool jit_enabled(const vdec_config* cfg);
bool jit_pick(const vdec_config* cfg, const uint8_t* labels, int* family, uint32_t* mode, char* name, size_t name_len);
const JitEntry* jit_get(const vdec_config* cfg, int family, uint32_t mode, char* err, size_t errlen);
const char* jit_name(const JitEntry* e);
int jit_family(const JitEntry* e);
int jit_launches(const JitEntry* e);
size_t jit_bytes(const JitEntry* e, const LaunchGeom& g);
int jit_run(const JitEntry* e, const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch,
            cudaStream_t st);
long jit_compile_check(const vdec_config* cfg, int family, uint32_t mode, char* err, size_t errlen);

// ---- generic table-driven kernels (vdec_generic.cu) ----------------------------------------
// One CTA per frame, path metrics in shared memory, labels from a device table; any K <= 15.
int generic_launch(const LaunchGeom& g, uint32_t metric, const uint32_t* bfly_labels_dev, const void* in,
                   uint8_t* out, int32_t* path_metric, uint32_t* surv, cudaStream_t st);
size_t generic_surv_bytes(const LaunchGeom& g);
size_t generic_smem_bytes(const LaunchGeom& g);

// ---- aux kernels (vdec_synth.cu) ---------------------------------------------------------------
int count_errors_launch(const uint8_t* dec, const uint8_t* payload, uint64_t nframes, uint32_t L,
                        uint32_t out_bytes, uint64_t* counters, cudaStream_t st);
int stitch_launch(const uint8_t* tmp, uint32_t row_bytes, uint32_t skip_bytes, uint64_t nblk, uint64_t keep_bytes,
                  uint64_t last_keep_bytes, uint8_t* out, cudaStream_t st);
int demap_launch(uint32_t n, const float* iq, uint64_t nsym, float* metrics, cudaStream_t st, bool hard = false);
int synth_awgn_launch(const LaunchGeom& g, const uint8_t* labels_dev, uint64_t nframes, uint64_t first_frame,
                      uint64_t seed, float sigma, uint8_t* payload, float* metrics, float* iq, cudaStream_t st);
int synth_bsc_launch(const LaunchGeom& g, const uint8_t* labels_dev, uint64_t nframes, uint64_t first_frame,
                     uint64_t seed, uint32_t crossover_ppm, uint8_t* payload, uint8_t* symbols, cudaStream_t st);

#endif  // !__CUDACC_RTC__

}  // namespace vd

#ifndef __CUDACC_RTC__
// The opaque handle of include/vdec.h
struct vdec {
    vdec_config cfg;
    vd::LaunchGeom geom;           // nframes filled per launch
    int device;
    uint32_t max_frames;             // frames per launch
    uint8_t labels_host[2 * 32768];  // [S][2], K <= 15 -> S <= 16384
    uint8_t* labels_dev;             // [S][2]
    uint32_t* bfly_labels_dev;       // [S/2] packed l00 | l01<<8 | l10<<16 | l11<<24
    const vd::RegsEntry* regs;     // non-null => specialised kernel
    const vd::Regs2Entry* regs2;   // non-null => packed 16-bit kernel (takes precedence)
    const vd::WideEntry* regsw;    // non-null => lane-group register kernel (K = 8 .. 12)
    const vd::TileEntry* tile;     // non-null => register-tile kernel (K = 8 .. 15, fp32 soft metrics)
    const vd::Tile2Entry* tile2;   // non-null => packed 16-bit register-tile kernel (K = 8 .. 15, integer metrics)
    const vd::JitEntry* jit;       // non-null => a kernel of one of these families compiled at run time for this code
    // host-buffer pipeline (vdec_decode_host): two stages, each with its own stream and device staging
    cudaStream_t streams[2];
    void* dev_in[2]; uint8_t* dev_out[2]; int32_t* dev_pm[2];
    uint64_t launches;
    char kernel_name[128];
};
#endif  // !__CUDACC_RTC__
