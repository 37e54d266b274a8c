// vdec_api.cu — the C ABI of include/vdec.h: handle lifecycle, kernel selection, chunked launches
// and the host-buffer pipeline.  No torch types, no CPU decode path: every compute entry point
// either launches CUDA kernels or fails with -1.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vdec_internal.h"

namespace vd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (getenv("VDEC_QUIET") == nullptr) fprintf(stderr, "[vdec] %s\n", g_err);   // callers of the legacy API ignore returns
}

static size_t in_bytes_per_frame(const vdec* h) {
    if (h->cfg.metric == VDEC_HARD_U16) return (size_t)h->geom.T;
    if (h->cfg.metric == VDEC_SOFT_U8) return (size_t)h->geom.T * h->geom.M;
    return (size_t)h->geom.T * h->geom.M * sizeof(float);
}

}  // namespace vd

// survivor scratch, one slot per concurrently running stream (0: decode_device, 1/2: host pipeline)
struct vdec_scratch { uint32_t* p[3]; size_t bytes[3]; };

// Guard zones (environment VDEC_GUARD=1 at vdec_create): every device buffer the library allocates for itself gets
// kGuardBytes of a known pattern on either side, and vdec_debug_check() counts the guard bytes that changed — a
// bounds check of the kernels' WRITES that needs no external tool (compute-sanitizer is not available on this pool).
static const size_t kGuardBytes = 4096;
static const unsigned char kGuardPattern = 0xA5;

// kept out of the public struct to keep vdec_internal.h small
struct vdec_priv {
    vdec base;
    vdec_scratch sc;
    size_t stage_frames;      // capacity of the host pipeline buffers, in frames
    float* met[3];            // demapped-metric staging for the raw-I/Q entry, per scratch slot
    size_t met_frames[3];
    uint8_t* stream_tmp[3];   // per-block outputs of the sliding-window decoder before stitching, per scratch slot
    size_t stream_tmp_bytes[3];
    cudaStream_t side[2];     // the first and the last window of a stream run beside the interior windows
    cudaEvent_t ev_fork, ev_join[2];
    float* iq_dev[2];         // device staging of host I/Q, per pipeline stage
    uint8_t* iq_out[2];       // and of the decoded output
    size_t iq_frames;
    const vd::JitEntry* jit_iq;   // run-time compiled register kernel with the raw-I/Q feeder (compiled on first use)
    bool jit_iq_tried;
    bool guard;                   // VDEC_GUARD: allocations carry guard zones
};

// cudaMalloc / cudaFree of the library's own buffers, with guard zones when asked for
static cudaError_t guarded_malloc(vdec_priv* p, void** out, size_t bytes);
static void guarded_free(vdec_priv* p, void* ptr);

static vdec_priv* priv(vdec_t* h) { return reinterpret_cast<vdec_priv*>(h); }

struct GuardRec { void* user; size_t bytes; };
static GuardRec g_guards[64];
static int g_nguards = 0;

static cudaError_t guarded_malloc(vdec_priv* p, void** out, size_t bytes) {
    if (!p->guard) return cudaMalloc(out, bytes);
    unsigned char* base = nullptr;
    const size_t padded = (bytes + 255) & ~(size_t)255;                // keep the tail guard aligned
    cudaError_t e = cudaMalloc(&base, padded + 2 * kGuardBytes);
    if (e != cudaSuccess) return e;
    cudaMemset(base, kGuardPattern, kGuardBytes);
    cudaMemset(base + kGuardBytes + bytes, kGuardPattern, padded - bytes + kGuardBytes);
    *out = base + kGuardBytes;
    if (g_nguards < 64) g_guards[g_nguards++] = GuardRec{*out, bytes};
    return cudaSuccess;
}
static void guarded_free(vdec_priv* p, void* ptr) {
    if (!ptr) return;
    if (!p->guard) { cudaFree(ptr); return; }
    for (int i = 0; i < g_nguards; ++i)
        if (g_guards[i].user == ptr) { g_guards[i] = g_guards[--g_nguards]; break; }
    cudaFree(static_cast<unsigned char*>(ptr) - kGuardBytes);
}

static int ensure_scratch(vdec_priv* p, int slot, const vd::LaunchGeom& geom, uint32_t frames) {
    vdec* h = &p->base;
    vd::LaunchGeom g = geom;
    g.nframes = frames;
    const size_t need = h->jit ? vd::jit_bytes(h->jit, g) : h->tile2 ? vd::tile2_bytes(h->tile2, g) : h->tile ? vd::tile_bytes(h->tile, g) : h->regsw ? vd::regsw_bytes(h->regsw, g) : h->regs2 ? vd::regs2_bytes(h->regs2, g)
                        : h->regs ? h->regs->surv_bytes(g) : vd::generic_surv_bytes(g);
    if (need <= p->sc.bytes[slot]) return 0;
    if (p->sc.p[slot]) { guarded_free(p, p->sc.p[slot]); p->sc.p[slot] = nullptr; p->sc.bytes[slot] = 0; }
    VDEC_CUDA(guarded_malloc(p, reinterpret_cast<void**>(&p->sc.p[slot]), need));
    p->sc.bytes[slot] = need;
    return 0;
}

static int decode_on(vdec_priv* p, int slot, const void* in_dev, uint64_t nframes, uint8_t* out_dev, int32_t* pm_dev,
                     cudaStream_t st);
static int decode_geom(vdec_priv* p, int slot, const vd::LaunchGeom& geom, const void* in_dev, uint64_t nframes,
                       uint8_t* out_dev, int32_t* pm_dev, cudaStream_t st, bool iq = false);

// does this handle decode raw I/Q inside its decoder kernel (metrics formed in the feeder, vdec_regs.cuh IqRing)?
static bool iq_fused(vdec_priv* p) {
    vdec* h = &p->base;
    if (h->cfg.metric != VDEC_SOFT_F32) return false;
    if (h->regs) return h->regs->launch_iq != nullptr;
    if (h->jit && vd::jit_family(h->jit) == vd::kJitRegs) {
        if (!p->jit_iq_tried) {
            p->jit_iq_tried = true;
            char why[2048] = "";
            p->jit_iq = vd::jit_get(&h->cfg, vd::kJitRegsIq, VDEC_SOFT_F32, why, sizeof(why));
            if (!p->jit_iq && why[0] && getenv("VDEC_QUIET") == nullptr)
                fprintf(stderr, "[vdec] raw-I/Q kernel not compiled, using the separate demapper: %s\n", why);
        }
        return p->jit_iq != nullptr;
    }
    return false;
}

// Raw I/Q in.  Register kernels (K <= 7) form the branch metrics in their feeder: one launch, 8 B of HBM per step.
// The other families go through the stand-alone demapper into a bounded staging buffer.
static int decode_iq_on(vdec_priv* p, int slot, const float* iq_dev, uint64_t nframes, uint8_t* out_dev, cudaStream_t st) {
    vdec* h = &p->base;
    if (h->cfg.metric != VDEC_SOFT_F32) { vd::set_error("raw-I/Q entry needs the soft metric"); return -1; }
    const vd::LaunchGeom& g = h->geom;
    if (iq_fused(p)) return decode_geom(p, slot, g, iq_dev, nframes, out_dev, nullptr, st, true);
    size_t chunk = ((size_t)1 << 30) / ((size_t)g.T * g.M * 4);      // ~1 GiB of staged metrics
    if (chunk < 1) chunk = 1;
    if (chunk > nframes) chunk = nframes;
    if (chunk > p->met_frames[slot]) {
        if (p->met[slot]) { guarded_free(p, p->met[slot]); p->met[slot] = nullptr; p->met_frames[slot] = 0; }
        VDEC_CUDA(guarded_malloc(p, reinterpret_cast<void**>(&p->met[slot]), chunk * g.T * g.M * sizeof(float)));
        p->met_frames[slot] = chunk;
    }
    for (uint64_t done = 0; done < nframes; done += chunk) {
        const uint64_t n = (nframes - done < chunk) ? (nframes - done) : chunk;
        if (vd::demap_launch(g.n, iq_dev + done * g.T * 2, n * g.T, p->met[slot], st) != 0) return -1;
        h->launches += 1;
        if (decode_on(p, slot, p->met[slot], n, out_dev + done * g.out_bytes, nullptr, st) != 0) return -1;
    }
    return 0;
}

// decode `nframes` frames of geometry `geom` (geom.T steps each, starts geom.stride_steps apart)
static int decode_geom(vdec_priv* p, int slot, const vd::LaunchGeom& geom, const void* in_dev, uint64_t nframes,
                       uint8_t* out_dev, int32_t* pm_dev, cudaStream_t st, bool iq) {
    vdec* h = &p->base;
    const size_t step_bytes = iq ? 2 * sizeof(float) : vd::in_bytes_per_frame(h) / h->geom.T;
    uint64_t done = 0;
    while (done < nframes) {
        const uint32_t n = (uint32_t)((nframes - done < h->max_frames) ? (nframes - done) : h->max_frames);
        if (ensure_scratch(p, slot, geom, n) != 0) return -1;
        vd::LaunchGeom g = geom;
        g.nframes = n;
        if (geom.init_all_from != 0xFFFFFFFFu)       // the index is relative to the first frame of the call
            g.init_all_from = geom.init_all_from > done ? (uint32_t)(geom.init_all_from - done) : 0u;
        const void* in = static_cast<const uint8_t*>(in_dev) + done * geom.stride_steps * step_bytes;
        uint8_t* out = out_dev + done * g.out_bytes;
        int32_t* pm = pm_dev ? pm_dev + done : nullptr;
        int rc;
        if (iq) rc = h->regs ? h->regs->launch_iq(g, in, out, pm, p->sc.p[slot], st) : vd::jit_run(p->jit_iq, g, in, out, pm, p->sc.p[slot], st);
        else if (h->jit) rc = vd::jit_run(h->jit, g, in, out, pm, p->sc.p[slot], st);
        else if (h->tile2) rc = vd::tile2_run(h->tile2, g, in, out, pm, p->sc.p[slot], st);
        else if (h->tile) rc = vd::tile_run(h->tile, g, in, out, p->sc.p[slot], st);
        else if (h->regsw) rc = vd::regsw_run(h->regsw, g, in, out, p->sc.p[slot], st);
        else if (h->regs2) rc = vd::regs2_run(h->regs2, g, in, out, pm, p->sc.p[slot], st);
        else if (h->regs) rc = h->regs->launch(g, in, out, pm, p->sc.p[slot], st);
        else rc = vd::generic_launch(g, h->cfg.metric, h->bfly_labels_dev, in, out, pm, p->sc.p[slot], st);
        if (rc != 0) return -1;
        h->launches += h->jit ? vd::jit_launches(h->jit)
                                : (!h->tile && !h->tile2 && (h->regs || h->regs2 || h->regsw)) ? 1 : 2;   // tile, table-driven = forward + traceback
        done += n;
    }
    return 0;
}

static int decode_on(vdec_priv* p, int slot, const void* in_dev, uint64_t nframes, uint8_t* out_dev, int32_t* pm_dev,
                     cudaStream_t st) {
    return decode_geom(p, slot, p->base.geom, in_dev, nframes, out_dev, pm_dev, st);
}


// 16-byte alignment of float inputs: the soft kernels stage rows with 16-byte cp.async / vector loads, and a misaligned
// device pointer would fault the context instead of returning -1
static bool aligned_input(const vdec* h, const void* in_dev, const char* who) {
    if (h->cfg.metric == VDEC_SOFT_F32 && (reinterpret_cast<uintptr_t>(in_dev) & 15u) != 0) {
        vd::set_error("%s: soft input must be 16-byte aligned", who);
        return false;
    }
    return true;
}

// packed 16-bit families: limits that depend on the number of steps of a launch
static bool packed_limits_ok(const vdec_config& cfg, uint64_t T) {
    const uint64_t mx = cfg.metric == VDEC_HARD_U16 ? cfg.symlen_out : (cfg.metric_max ? cfg.metric_max : 255);
    if (T * mx >= 0x4000u) return false;
    return cfg.constr_len >= 8 || T * 32 <= 200 * 1024;          // register kernels: output tile in shared memory
}

// vdec_stream_decode_device: swap the handle's kernel family for the duration of one call when the window geometry
// breaks the limits of the selected one (restored on scope exit; a handle has one call in flight at a time)
struct StreamSelGuard {
    vdec* h;
    const vd::RegsEntry* regs; const vd::Regs2Entry* regs2; const vd::Tile2Entry* tile2; const vd::JitEntry* jit;
    bool swapped;
    StreamSelGuard(vdec* h_, uint32_t window_T) : h(h_), regs(h_->regs), regs2(h_->regs2), tile2(h_->tile2), jit(h_->jit), swapped(false) {
        const bool packed = h->regs2 || h->tile2 ||
                            (h->jit && (vd::jit_family(h->jit) == vd::kJitRegs2 || vd::jit_family(h->jit) == vd::kJitTile2));
        const bool regs_too_long = (h->regs || (h->jit && vd::jit_family(h->jit) == vd::kJitRegs)) && (uint64_t)window_T * 16 > 200 * 1024;
        if ((packed && !packed_limits_ok(h->cfg, window_T)) || regs_too_long) {
            swapped = true;
            h->regs2 = nullptr; h->tile2 = nullptr; h->jit = nullptr; h->regs = nullptr;
            if (!regs_too_long && h->cfg.metric == VDEC_HARD_U16 && h->cfg.constr_len <= 7 && (uint64_t)window_T * 16 <= 200 * 1024)
                h->regs = vd::regs_find(h->cfg.constr_len, h->cfg.symlen_out, VDEC_HARD_U16, h->labels_host);
        }
    }
    ~StreamSelGuard() { if (swapped) { h->regs = regs; h->regs2 = regs2; h->tile2 = tile2; h->jit = jit; } }
};

static const char* metric_tag(uint32_t metric) {
    return metric == VDEC_HARD_U16 ? "hard_u16" : metric == VDEC_SOFT_U8 ? "soft_u8" : "soft_f32";
}

// kernel family for a config: packed 16-bit (integer metrics that provably stay below 2^15) >
// specialised register kernel > table-driven
static void select_kernels(const vdec_config* cfg, const uint8_t* labels, const vd::RegsEntry** regs,
                           const vd::Regs2Entry** regs2, const vd::WideEntry** regsw, const vd::TileEntry** tile,
                           const vd::Tile2Entry** tile2) {
    *regs = nullptr; *regs2 = nullptr; *regsw = nullptr; *tile = nullptr; *tile2 = nullptr;
    if (cfg->flags & VDEC_FORCE_GENERIC) return;
    if (cfg->constr_len >= 8) {
        if (cfg->metric != VDEC_SOFT_F32) {          // integer metrics: packed 16-bit tile kernel when the range allows
            const uint64_t Ti = (uint64_t)cfg->block_len + cfg->constr_len - 1;
            const uint64_t mx = cfg->metric == VDEC_HARD_U16 ? cfg->symlen_out : (cfg->metric_max ? cfg->metric_max : 255);
            if (Ti * mx < 0x4000u)
                *tile2 = vd::tile2_find(cfg->constr_len, cfg->symlen_out, cfg->metric == VDEC_HARD_U16 ? 0 : 1, labels);
            return;
        }
        if (!(cfg->flags & VDEC_PREFER_LANEGROUP) || cfg->constr_len > 12)
            *tile = vd::tile_find(cfg->constr_len, cfg->symlen_out, cfg->metric, labels);
        if (!*tile && (uint64_t)cfg->block_len + cfg->constr_len - 1 <= 65536)
            *regsw = vd::regsw_find(cfg->constr_len, cfg->symlen_out, cfg->metric, labels);
        return;
    }
    const uint64_t T = (uint64_t)cfg->block_len + cfg->constr_len - 1;
    // the register kernels stage a warp's output rows in shared memory: 16 B (32 B packed) per step and CTA
    const bool fits1 = T * 16 <= 200 * 1024, fits2 = T * 32 <= 200 * 1024;
    if (!fits1) return;
    const bool small = T * (cfg->metric == VDEC_HARD_U16 ? cfg->symlen_out : (cfg->metric_max ? cfg->metric_max : 255)) < 0x4000u;
    if (cfg->metric == VDEC_SOFT_U8) {
        if (fits2 && small) *regs2 = vd::regs2_find(cfg->constr_len, cfg->symlen_out, 1, labels);
        return;
    }
    // hard decisions: packed 16-bit kernel (2 issue slots per state update; 128 vs 75 Gbit/s at K=7, 627 vs 467
    // at K=3, tools/bench_configs.py) when the metric range allows, else the 32-bit register kernel
    if (!(cfg->flags & VDEC_PREFER_32BIT) && cfg->metric == VDEC_HARD_U16 && fits2 && small)
        *regs2 = vd::regs2_find(cfg->constr_len, cfg->symlen_out, 0, labels);
    if (*regs2) return;
    *regs = vd::regs_find(cfg->constr_len, cfg->symlen_out, cfg->metric, labels);
    if (!*regs && cfg->metric == VDEC_HARD_U16 && fits2 && small)
        *regs2 = vd::regs2_find(cfg->constr_len, cfg->symlen_out, 0, labels);
}

extern "C" {

const char* vdec_last_error(void) { return vd::g_err; }

int vdec_create(vdec_t** out, const struct vdec_config* cfg) {
    if (!out || !cfg) { vd::set_error("vdec_create: null argument"); return -1; }
    *out = nullptr;
    if (cfg->symlen_out < 1 || cfg->symlen_out > 3) { vd::set_error("symlen_out %u unsupported (1..3)", cfg->symlen_out); return -1; }
    if (cfg->constr_len < 2 || cfg->constr_len > 15) { vd::set_error("constr_len %u unsupported (2..15)", cfg->constr_len); return -1; }
    if (cfg->block_len < 1 || cfg->block_len > (1u << 24)) { vd::set_error("block_len %u unsupported", cfg->block_len); return -1; }
    if (cfg->metric > VDEC_SOFT_U8 || cfg->parity > VDEC_PARITY_LINEAR) { vd::set_error("bad metric/parity selector"); return -1; }
    if (cfg->metric == VDEC_SOFT_U8) {
        // the integer kernels sum exactly; the reference sums the same integers in fp32, which is exact (hence equal) only
        // while every path metric stays below 2^24
        const uint64_t Tq = (uint64_t)cfg->block_len + cfg->constr_len - 1, mx = cfg->metric_max ? cfg->metric_max : 255;
        if (Tq * mx >= (1ull << 24)) {
            vd::set_error("VDEC_SOFT_U8: (block_len + constr_len - 1) * metric_max = %llu reaches 2^24, where the reference's fp32 sums "
                          "stop being exact; use VDEC_SOFT_F32", (unsigned long long)(Tq * mx));
            return -1;
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        vd::set_error("no CUDA device: this library has no CPU path"); return -1;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { vd::set_error("device %d out of range (%d present)", cfg->device, ndev); return -1; }
    VDEC_CUDA(cudaSetDevice(cfg->device));

    vdec_priv* p = static_cast<vdec_priv*>(calloc(1, sizeof(vdec_priv)));
    if (!p) { vd::set_error("out of host memory"); return -1; }
    vdec* h = &p->base;
    h->cfg = *cfg;
    h->device = cfg->device;
    { const char* g = getenv("VDEC_GUARD"); p->guard = g && *g && *g != '0'; }
    vd::LaunchGeom& g = h->geom;
    g.K = cfg->constr_len; g.n = cfg->symlen_out; g.L = cfg->block_len;
    g.S = 1u << (g.K - 1); g.T = g.L + g.K - 1; g.M = 1u << g.n;
    g.out_bytes = (g.T + 7) / 8; g.nw = (g.S + 31) / 32; g.nframes = 0;
    g.stride_steps = g.T; g.init_all_from = 0xFFFFFFFFu;

    // trellis labels from the caller's polynomials (read now: the pointer is only borrowed, viterbi-decoder.c:169)
    const bool quirk = cfg->parity == VDEC_PARITY_REF;
    for (uint32_t s = 0; s < g.S; ++s)
        for (uint32_t i = 0; i < 2; ++i)
            h->labels_host[2 * s + i] = (uint8_t)vd::branch_label(g.K, g.n, cfg->polynomials, quirk, s, i);
    uint32_t* bl = static_cast<uint32_t*>(malloc(sizeof(uint32_t) * (g.S / 2)));
    if (!bl) { free(p); vd::set_error("out of host memory"); return -1; }
    for (uint32_t m = 0; m < g.S / 2; ++m)
        bl[m] = (uint32_t)h->labels_host[2 * (2 * m)] | ((uint32_t)h->labels_host[2 * (2 * m + 1)] << 8) |
                ((uint32_t)h->labels_host[2 * (2 * m) + 1] << 16) | ((uint32_t)h->labels_host[2 * (2 * m + 1) + 1] << 24);
    cudaError_t e = cudaMalloc(&h->labels_dev, 2 * g.S);
    if (e == cudaSuccess) e = cudaMalloc(&h->bfly_labels_dev, sizeof(uint32_t) * (g.S / 2));
    if (e == cudaSuccess) e = cudaMemcpy(h->labels_dev, h->labels_host, 2 * g.S, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->bfly_labels_dev, bl, sizeof(uint32_t) * (g.S / 2), cudaMemcpyHostToDevice);
    free(bl);
    if (e != cudaSuccess) { vd::set_error("label upload failed: %s", cudaGetErrorString(e)); vdec_destroy(h); return -1; }

    select_kernels(cfg, h->labels_host, &h->regs, &h->regs2, &h->regsw, &h->tile, &h->tile2);
    int jit_family = 0; uint32_t jit_mode = 0;
    if (!h->regs && !h->regs2 && !h->regsw && !h->tile && !h->tile2 &&
        vd::jit_pick(cfg, h->labels_host, &jit_family, &jit_mode, nullptr, 0)) {
        // polynomials outside the registries: compile the matching kernel family for them (vdec_jit.cu)
        char why[2048] = "";
        h->jit = vd::jit_get(cfg, jit_family, jit_mode, why, sizeof(why));
        if (!h->jit && why[0] && getenv("VDEC_QUIET") == nullptr)
            fprintf(stderr, "[vdec] run-time compilation unavailable, using the table-driven kernels: %s\n", why);
    }
    const bool jit_tile = h->jit && vd::jit_launches(h->jit) == 2;
    if (h->tile || h->tile2 || jit_tile) {
        snprintf(h->kernel_name, sizeof(h->kernel_name), "%s", h->jit ? vd::jit_name(h->jit) : h->tile2 ? vd::tile2_name(h->tile2) : vd::tile_name(h->tile));
        // bound the survivor scratch of one launch to ~8 GiB (S/8 bytes per step and frame)
        // (the packed kernels store the decisions of two frames in the same words: same bytes per frame)
        size_t cap = ((size_t)8 << 30) / ((size_t)g.T * (g.S / 8));
        if (cap < 1) cap = 1;
        if (cap > (1u << 20)) cap = 1u << 20;
        h->max_frames = cfg->max_frames ? cfg->max_frames : (uint32_t)cap;
        if (h->max_frames > cap) h->max_frames = (uint32_t)cap;
    } else if (h->regs2 || h->regs || h->regsw || h->jit) {
        snprintf(h->kernel_name, sizeof(h->kernel_name), "%s",
                 h->jit ? vd::jit_name(h->jit) : h->regsw ? vd::regsw_name(h->regsw) : h->regs2 ? vd::regs2_name(h->regs2) : h->regs->name);
        h->max_frames = cfg->max_frames ? cfg->max_frames : (1u << 20);
    } else {
        snprintf(h->kernel_name, sizeof(h->kernel_name), "generic_%s<K%u n%u table-driven>",
                 metric_tag(cfg->metric), g.K, g.n);
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
        if (vd::generic_smem_bytes(g) > (size_t)max_smem) {
            vd::set_error("K=%u needs %zu B shared memory per CTA, device offers %d", g.K, vd::generic_smem_bytes(g), max_smem);
            vdec_destroy(h); return -1;
        }
        // bound the survivor scratch of one launch to ~4 GiB
        const size_t per_frame = (size_t)g.T * g.nw * 4;
        size_t cap = ((size_t)4 << 30) / per_frame;
        if (cap < 1) cap = 1;
        if (cap > (1u << 20)) cap = 1u << 20;
        h->max_frames = cfg->max_frames ? cfg->max_frames : (uint32_t)cap;
        if (h->max_frames > cap) h->max_frames = (uint32_t)cap;
    }
    *out = h;
    return 0;
}

void vdec_destroy(vdec_t* h) {
    if (!h) return;
    vdec_priv* p = priv(h);
    cudaSetDevice(h->device);
    for (int i = 0; i < 3; ++i) guarded_free(p, p->sc.p[i]);
    for (int i = 0; i < 3; ++i) guarded_free(p, p->met[i]);
    for (int i = 0; i < 3; ++i) guarded_free(p, p->stream_tmp[i]);
    for (int i = 0; i < 2; ++i) { if (p->side[i]) cudaStreamDestroy(p->side[i]); if (p->ev_join[i]) cudaEventDestroy(p->ev_join[i]); }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    for (int i = 0; i < 2; ++i) { if (p->iq_dev[i]) cudaFree(p->iq_dev[i]); if (p->iq_out[i]) cudaFree(p->iq_out[i]); }
    for (int i = 0; i < 2; ++i) {
        if (h->dev_in[i]) cudaFree(h->dev_in[i]);
        if (h->dev_out[i]) cudaFree(h->dev_out[i]);
        if (h->dev_pm[i]) cudaFree(h->dev_pm[i]);
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    if (h->labels_dev) cudaFree(h->labels_dev);
    if (h->bfly_labels_dev) cudaFree(h->bfly_labels_dev);
    free(p);
}

uint32_t vdec_steps(const vdec_t* h) { return h->geom.T; }
uint32_t vdec_states(const vdec_t* h) { return h->geom.S; }
size_t vdec_in_bytes_per_frame(const vdec_t* h) { return vd::in_bytes_per_frame(h); }
size_t vdec_out_bytes_per_frame(const vdec_t* h) { return h->geom.out_bytes; }
const char* vdec_kernel_name(const vdec_t* h) { return h->kernel_name; }
uint64_t vdec_launch_count(const vdec_t* h) { return h->launches; }

int vdec_config_labels(const struct vdec_config* cfg, uint8_t* labels) {
    if (!cfg || !labels) { vd::set_error("vdec_config_labels: null argument"); return -1; }
    if (cfg->symlen_out < 1 || cfg->symlen_out > 3 || cfg->constr_len < 2 || cfg->constr_len > 15) {
        vd::set_error("vdec_config_labels: unsupported code"); return -1;
    }
    const uint32_t S = 1u << (cfg->constr_len - 1);
    for (uint32_t s = 0; s < S; ++s)
        for (uint32_t i = 0; i < 2; ++i)
            labels[2 * s + i] = (uint8_t)vd::branch_label(cfg->constr_len, cfg->symlen_out, cfg->polynomials,
                                                          cfg->parity == VDEC_PARITY_REF, s, i);
    return 0;
}

const char* vdec_config_kernel_name(const struct vdec_config* cfg) {
    static thread_local char name[128];
    uint8_t labels[2 * 16384];
    if (vdec_config_labels(cfg, labels) != 0) return "";
    const vd::RegsEntry* e = nullptr; const vd::Regs2Entry* e2 = nullptr; const vd::WideEntry* ew = nullptr;
    const vd::TileEntry* et = nullptr; const vd::Tile2Entry* et2 = nullptr;
    select_kernels(cfg, labels, &e, &e2, &ew, &et, &et2);
    if (et2) snprintf(name, sizeof(name), "%s", vd::tile2_name(et2));
    else if (et) snprintf(name, sizeof(name), "%s", vd::tile_name(et));
    else if (ew) snprintf(name, sizeof(name), "%s", vd::regsw_name(ew));
    else if (e2) snprintf(name, sizeof(name), "%s", vd::regs2_name(e2));
    else if (e) snprintf(name, sizeof(name), "%s", e->name);
    else {
        int family = 0; uint32_t mode = 0;
        if (!vd::jit_pick(cfg, labels, &family, &mode, name, sizeof(name)))
            snprintf(name, sizeof(name), "generic_%s<K%u n%u table-driven>", metric_tag(cfg->metric), cfg->constr_len, cfg->symlen_out);
    }
    return name;
}

long vdec_jit_check(const struct vdec_config* cfg) {
    static thread_local uint8_t labels[2 * 16384];
    if (vdec_config_labels(cfg, labels) != 0) return -1;
    const vd::RegsEntry* e = nullptr; const vd::Regs2Entry* e2 = nullptr; const vd::WideEntry* ew = nullptr;
    const vd::TileEntry* et = nullptr; const vd::Tile2Entry* et2 = nullptr;
    select_kernels(cfg, labels, &e, &e2, &ew, &et, &et2);
    int family = 0; uint32_t mode = 0;
    if (e || e2 || ew || et || et2 || !vd::jit_pick(cfg, labels, &family, &mode, nullptr, 0)) return 0;
    char why[2048] = "";
    const long n = vd::jit_compile_check(cfg, family, mode, why, sizeof(why));
    if (n < 0) vd::set_error("run-time compilation failed: %s", why);
    return n;
}

long vdec_jit_check_iq(const struct vdec_config* cfg) {
    static thread_local uint8_t labels[2 * 16384];
    if (vdec_config_labels(cfg, labels) != 0) return -1;
    int family = 0; uint32_t mode = 0;
    if (cfg->metric != VDEC_SOFT_F32 || !vd::jit_pick(cfg, labels, &family, &mode, nullptr, 0) || family != vd::kJitRegs) return 0;
    char why[2048] = "";
    const long n = vd::jit_compile_check(cfg, vd::kJitRegsIq, VDEC_SOFT_F32, why, sizeof(why));
    if (n < 0) vd::set_error("run-time compilation failed: %s", why);
    return n;
}

int vdec_labels(const vdec_t* h, uint8_t* labels) {
    if (!h || !labels) { vd::set_error("vdec_labels: null argument"); return -1; }
    memcpy(labels, h->labels_host, 2 * h->geom.S);
    return 0;
}

int vdec_decode_device(vdec_t* h, const void* in_dev, uint64_t nframes, uint8_t* out_dev, int32_t* path_metric_dev,
                       void* cuda_stream) {
    if (!h || (nframes && (!in_dev || !out_dev))) { vd::set_error("vdec_decode_device: null argument"); return -1; }
    if (nframes == 0) return 0;
    if (!aligned_input(h, in_dev, "vdec_decode_device")) return -1;
    VDEC_CUDA(cudaSetDevice(h->device));
    return decode_on(priv(h), 0, in_dev, nframes, out_dev, path_metric_dev, static_cast<cudaStream_t>(cuda_stream));
}

int vdec_decode_host(vdec_t* h, const void* in_host, uint64_t nframes, uint8_t* out_host, int32_t* path_metric_host) {
    if (!h || (nframes && (!in_host || !out_host))) { vd::set_error("vdec_decode_host: null argument"); return -1; }
    if (nframes == 0) return 0;
    vdec_priv* p = priv(h);
    VDEC_CUDA(cudaSetDevice(h->device));
    const size_t ib = vd::in_bytes_per_frame(h), ob = h->geom.out_bytes;
    // pipeline stage size: ~64 MiB of input, at least 1 frame, at most what is asked for
    size_t chunk = ((size_t)64 << 20) / ib;
    if (chunk < 1) chunk = 1;
    if (chunk > h->max_frames) chunk = h->max_frames;
    if (chunk > nframes) chunk = nframes;
    if (chunk > p->stage_frames) {
        for (int i = 0; i < 2; ++i) {
            if (h->dev_in[i]) { cudaFree(h->dev_in[i]); h->dev_in[i] = nullptr; }
            if (h->dev_out[i]) { cudaFree(h->dev_out[i]); h->dev_out[i] = nullptr; }
            if (h->dev_pm[i]) { cudaFree(h->dev_pm[i]); h->dev_pm[i] = nullptr; }
        }
        p->stage_frames = 0;
        for (int i = 0; i < 2; ++i) {
            if (!h->streams[i]) VDEC_CUDA(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
            VDEC_CUDA(cudaMalloc(&h->dev_in[i], chunk * ib));
            VDEC_CUDA(cudaMalloc(&h->dev_out[i], chunk * ob));
            VDEC_CUDA(cudaMalloc(&h->dev_pm[i], chunk * sizeof(int32_t)));
        }
        p->stage_frames = chunk;
    }
    const bool want_pm = path_metric_host != nullptr && h->cfg.metric == VDEC_HARD_U16;
    uint64_t done = 0;
    int stage = 0;
    while (done < nframes) {
        const uint64_t n = (nframes - done < chunk) ? (nframes - done) : chunk;
        cudaStream_t st = h->streams[stage];
        // the stage's previous D2H must have drained before its buffers are reused: stream order does that
        VDEC_CUDA(cudaMemcpyAsync(h->dev_in[stage], static_cast<const uint8_t*>(in_host) + done * ib, n * ib,
                                  cudaMemcpyHostToDevice, st));
        if (decode_on(p, 1 + stage, h->dev_in[stage], n, h->dev_out[stage], want_pm ? h->dev_pm[stage] : nullptr, st) != 0) return -1;
        VDEC_CUDA(cudaMemcpyAsync(out_host + done * ob, h->dev_out[stage], n * ob, cudaMemcpyDeviceToHost, st));
        if (want_pm)
            VDEC_CUDA(cudaMemcpyAsync(path_metric_host + done, h->dev_pm[stage], n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        done += n;
        stage ^= 1;
    }
    VDEC_CUDA(cudaStreamSynchronize(h->streams[0]));
    VDEC_CUDA(cudaStreamSynchronize(h->streams[1]));
    if (path_metric_host != nullptr && !want_pm) memset(path_metric_host, 0, nframes * sizeof(int32_t));
    return 0;
}

int vdec_decode_iq_device(vdec_t* h, const float* iq_dev, uint64_t nframes, uint8_t* out_dev, void* cuda_stream) {
    if (!h || (nframes && (!iq_dev || !out_dev))) { vd::set_error("vdec_decode_iq_device: null argument"); return -1; }
    if (nframes == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(iq_dev) & 15u) != 0) { vd::set_error("vdec_decode_iq_device: iq_dev must be 16-byte aligned"); return -1; }
    VDEC_CUDA(cudaSetDevice(h->device));
    return decode_iq_on(priv(h), 0, iq_dev, nframes, out_dev, static_cast<cudaStream_t>(cuda_stream));
}

int vdec_decode_iq_host(vdec_t* h, const float* iq_host, uint64_t nframes, uint8_t* out_host) {
    if (!h || (nframes && (!iq_host || !out_host))) { vd::set_error("vdec_decode_iq_host: null argument"); return -1; }
    if (nframes == 0) return 0;
    vdec_priv* p = priv(h);
    VDEC_CUDA(cudaSetDevice(h->device));
    const size_t ib = (size_t)h->geom.T * 2 * sizeof(float), ob = h->geom.out_bytes;
    size_t chunk = ((size_t)64 << 20) / ib;
    if (chunk < 1) chunk = 1;
    if (chunk > h->max_frames) chunk = h->max_frames;
    if (chunk > nframes) chunk = nframes;
    if (chunk > p->iq_frames) {
        for (int i = 0; i < 2; ++i) {
            if (p->iq_dev[i]) { cudaFree(p->iq_dev[i]); p->iq_dev[i] = nullptr; }
            if (p->iq_out[i]) { cudaFree(p->iq_out[i]); p->iq_out[i] = nullptr; }
        }
        p->iq_frames = 0;
        for (int i = 0; i < 2; ++i) {
            VDEC_CUDA(cudaMalloc(&p->iq_dev[i], chunk * ib));
            VDEC_CUDA(cudaMalloc(&p->iq_out[i], chunk * ob));
        }
        p->iq_frames = chunk;
    }
    for (int i = 0; i < 2; ++i)
        if (!h->streams[i]) VDEC_CUDA(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
    uint64_t done = 0;
    int stage = 0;
    while (done < nframes) {
        const uint64_t n = (nframes - done < chunk) ? (nframes - done) : chunk;
        cudaStream_t st = h->streams[stage];
        VDEC_CUDA(cudaMemcpyAsync(p->iq_dev[stage], iq_host + done * h->geom.T * 2, n * ib, cudaMemcpyHostToDevice, st));
        if (decode_iq_on(p, 1 + stage, p->iq_dev[stage], n, p->iq_out[stage], st) != 0) return -1;
        VDEC_CUDA(cudaMemcpyAsync(out_host + done * ob, p->iq_out[stage], n * ob, cudaMemcpyDeviceToHost, st));
        done += n;
        stage ^= 1;
    }
    VDEC_CUDA(cudaStreamSynchronize(h->streams[0]));
    VDEC_CUDA(cudaStreamSynchronize(h->streams[1]));
    return 0;
}

// One window launch of the sliding-window decoder: `nblk` blocks of T steps whose starts are `stride` steps apart,
// decoded into the temporary buffer, then bytes [skip/8, skip/8 + keep_bytes) of each block are appended to `out`.
static int stream_pass(vdec_priv* p, int slot, const uint8_t* in, size_t step_bytes, uint64_t start_step, uint64_t nblk, uint32_t T,
                       uint32_t stride, bool known_first, uint32_t skip_bits, uint64_t keep_bytes, uint64_t last_keep_bytes,
                       uint8_t* out, cudaStream_t st) {
    vdec* h = &p->base;
    if (nblk == 0) return 0;
    vd::LaunchGeom g = h->geom;
    g.T = T; g.L = T - (g.K - 1 < T ? g.K - 1 : 0); g.out_bytes = (T + 7) / 8;
    g.stride_steps = stride; g.init_all_from = known_first ? 1u : 0u;
    const size_t need = nblk * g.out_bytes;
    if (need > p->stream_tmp_bytes[slot]) {
        if (p->stream_tmp[slot]) { guarded_free(p, p->stream_tmp[slot]); p->stream_tmp[slot] = nullptr; p->stream_tmp_bytes[slot] = 0; }
        VDEC_CUDA(guarded_malloc(p, reinterpret_cast<void**>(&p->stream_tmp[slot]), need));
        p->stream_tmp_bytes[slot] = need;
    }
    if (decode_geom(p, slot, g, in + start_step * step_bytes, nblk, p->stream_tmp[slot], nullptr, st) != 0) return -1;
    h->launches += 1;
    return vd::stitch_launch(p->stream_tmp[slot], g.out_bytes, skip_bits / 8, nblk, keep_bytes, last_keep_bytes, out, st);
}

int vdec_stream_decode_device(vdec_t* h, const void* in_dev, uint64_t nsteps, uint32_t block_steps, uint32_t depth,
                              uint32_t flags, uint8_t* out_dev, void* cuda_stream) {
    if (!h || !in_dev || !out_dev) { vd::set_error("vdec_stream_decode_device: null argument"); return -1; }
    if (block_steps < 8 || (block_steps % 8) || (depth % 8)) { vd::set_error("block_steps and depth must be multiples of 8"); return -1; }
    if (depth > block_steps) { vd::set_error("depth must not exceed block_steps"); return -1; }
    if ((uint64_t)block_steps + 2ull * depth > 8192) { vd::set_error("block_steps + 2*depth must not exceed 8192"); return -1; }
    if (nsteps == 0) return 0;
    if (!aligned_input(h, in_dev, "vdec_stream_decode_device")) return -1;
    vdec_priv* p = priv(h);
    // The kernel family was chosen for frames of block_len steps; the windows are block_steps + 2*depth long.  The packed
    // 16-bit families have no renormalisation (T * metric_max < 2^14) and stage 32 B per step in shared memory: when
    // the window breaks either limit this call runs on the 32-bit register kernel or the table-driven kernels instead.
    StreamSelGuard guard(h, (uint32_t)((uint64_t)block_steps + 2ull * depth < nsteps ? block_steps + 2 * depth : nsteps));
    VDEC_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const uint8_t* in = static_cast<const uint8_t*>(in_dev);
    const size_t step_bytes = vd::in_bytes_per_frame(h) / h->geom.T;
    const bool known = (flags & VDEC_STREAM_START_KNOWN) != 0;
    const uint64_t Lb = block_steps, D = depth;
    if (nsteps <= Lb + 2 * D) {                                  // short stream: one window, exact
        return stream_pass(p, 0, in, step_bytes, 0, 1, (uint32_t)nsteps, (uint32_t)nsteps, known, 0, 0, (nsteps + 7) / 8, out_dev, st);
    }
    const uint64_t nfull = (nsteps - D) / Lb;                    // windows that deliver a full block of Lb bits
    // The first and the last window are one frame each (a single warp walking > 1000 steps: 0.3 ms of pure latency);
    // they run on two side streams of the handle, forked from and joined back into the caller's stream, beside the
    // launch of all interior windows, each with its own scratch slot and staging buffer.
    for (int i = 0; i < 2; ++i) {
        if (!p->side[i]) VDEC_CUDA(cudaStreamCreateWithFlags(&p->side[i], cudaStreamNonBlocking));
        if (!p->ev_join[i]) VDEC_CUDA(cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming));
    }
    if (!p->ev_fork) VDEC_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    VDEC_CUDA(cudaEventRecord(p->ev_fork, st));
    for (int i = 0; i < 2; ++i) VDEC_CUDA(cudaStreamWaitEvent(p->side[i], p->ev_fork, 0));
    // first window: starts with the stream (no warm-up), D steps of look-ahead
    if (stream_pass(p, 1, in, step_bytes, 0, 1, (uint32_t)(Lb + D), (uint32_t)Lb, known, 0, 0, Lb / 8, out_dev, p->side[0]) != 0) return -1;
    // last window: runs to the end of the stream, traced back from the best final state (viterbi-decoder.c:72-79)
    const uint64_t s0 = nfull * Lb - D, rest_bits = nsteps - nfull * Lb;
    if (stream_pass(p, 2, in, step_bytes, s0, 1, (uint32_t)(nsteps - s0), (uint32_t)(nsteps - s0), false, (uint32_t)D, 0,
                    (rest_bits + 7) / 8, out_dev + nfull * (Lb / 8), p->side[1]) != 0) return -1;
    // interior windows: D steps warm-up (start state unknown), Lb payload, D look-ahead; starts Lb apart
    if (nfull > 1 && stream_pass(p, 0, in, step_bytes, Lb - D, nfull - 1, (uint32_t)(Lb + 2 * D), (uint32_t)Lb, false, (uint32_t)D,
                                 Lb / 8, Lb / 8, out_dev + Lb / 8, st) != 0) return -1;
    for (int i = 0; i < 2; ++i) {
        VDEC_CUDA(cudaEventRecord(p->ev_join[i], p->side[i]));
        VDEC_CUDA(cudaStreamWaitEvent(st, p->ev_join[i], 0));
    }
    return 0;
}

long vdec_debug_check(vdec_t* h) {
    if (!h) return -1;
    vdec_priv* p = priv(h);
    if (!p->guard) { vd::set_error("vdec_debug_check: the handle was not created with VDEC_GUARD=1"); return -1; }
    VDEC_CUDA(cudaSetDevice(h->device));
    VDEC_CUDA(cudaDeviceSynchronize());
    long bad = 0;
    unsigned char* host = static_cast<unsigned char*>(malloc(kGuardBytes + 256));
    if (!host) return -1;
    void* bufs[] = {p->sc.p[0], p->sc.p[1], p->sc.p[2], p->met[0], p->met[1], p->met[2], p->stream_tmp[0], p->stream_tmp[1], p->stream_tmp[2]};
    for (void* u : bufs) {
        if (!u) continue;
        size_t bytes = 0;
        for (int i = 0; i < g_nguards; ++i) if (g_guards[i].user == u) bytes = g_guards[i].bytes;
        const size_t padded = (bytes + 255) & ~(size_t)255;
        unsigned char* base = static_cast<unsigned char*>(u) - kGuardBytes;
        if (cudaMemcpy(host, base, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) { free(host); return -1; }
        for (size_t i = 0; i < kGuardBytes; ++i) bad += host[i] != kGuardPattern;
        const size_t tail = padded - bytes + kGuardBytes;
        if (cudaMemcpy(host, base + kGuardBytes + bytes, tail, cudaMemcpyDeviceToHost) != cudaSuccess) { free(host); return -1; }
        for (size_t i = 0; i < tail; ++i) bad += host[i] != kGuardPattern;
    }
    free(host);
    return bad;
}

int vdec_sync(vdec_t* h) {
    if (!h) return -1;
    VDEC_CUDA(cudaSetDevice(h->device));
    VDEC_CUDA(cudaDeviceSynchronize());
    return 0;
}

int vdec_count_errors_device(vdec_t* h, const uint8_t* decoded_dev, const uint8_t* payload_dev, uint64_t nframes,
                             uint64_t* counters_dev, void* cuda_stream) {
    if (!h || !decoded_dev || !payload_dev || !counters_dev) { vd::set_error("vdec_count_errors_device: null argument"); return -1; }
    VDEC_CUDA(cudaSetDevice(h->device));
    h->launches += 1;
    return vd::count_errors_launch(decoded_dev, payload_dev, nframes, h->geom.L, h->geom.out_bytes, counters_dev,
                                   static_cast<cudaStream_t>(cuda_stream));
}

int vdec_demap_device(uint32_t symlen_out, const float* iq_dev, uint64_t nsym, float* metrics_dev, void* cuda_stream) {
    if (!iq_dev || !metrics_dev) { vd::set_error("vdec_demap_device: null argument"); return -1; }
    return vd::demap_launch(symlen_out, iq_dev, nsym, metrics_dev, static_cast<cudaStream_t>(cuda_stream));
}

int vdec_hard_demap_device(uint32_t symlen_out, const float* iq_dev, uint64_t nsym, float* metrics_dev, void* cuda_stream) {
    if (!iq_dev || !metrics_dev) { vd::set_error("vdec_hard_demap_device: null argument"); return -1; }
    return vd::demap_launch(symlen_out, iq_dev, nsym, metrics_dev, static_cast<cudaStream_t>(cuda_stream), true);
}

const char* vdec_iq_kernel_name(vdec_t* h) {
    if (!h || h->cfg.metric != VDEC_SOFT_F32) return "";
    vdec_priv* p = priv(h);
    static thread_local char name[192];
    if (iq_fused(p)) snprintf(name, sizeof(name), "%s + demapper in the feeder (raw I/Q ring)", h->regs ? h->regs->name : vd::jit_name(p->jit_iq));
    else snprintf(name, sizeof(name), "demap_kernel -> %s", h->kernel_name);
    return name;
}

int vdec_synth_awgn_device(vdec_t* h, uint64_t nframes, uint64_t first_frame, uint64_t seed, float sigma,
                           uint8_t* payload_dev, float* metrics_dev, float* iq_dev, void* cuda_stream) {
    if (!h || !payload_dev || (!metrics_dev && !iq_dev)) { vd::set_error("vdec_synth_awgn_device: null argument"); return -1; }
    VDEC_CUDA(cudaSetDevice(h->device));
    h->launches += 2;
    return vd::synth_awgn_launch(h->geom, h->labels_dev, nframes, first_frame, seed, sigma, payload_dev, metrics_dev, iq_dev,
                                 static_cast<cudaStream_t>(cuda_stream));
}

int vdec_synth_bsc_device(vdec_t* h, uint64_t nframes, uint64_t first_frame, uint64_t seed, uint32_t crossover_ppm,
                          uint8_t* payload_dev, uint8_t* symbols_dev, void* cuda_stream) {
    if (!h || !payload_dev || !symbols_dev) { vd::set_error("vdec_synth_bsc_device: null argument"); return -1; }
    VDEC_CUDA(cudaSetDevice(h->device));
    h->launches += 2;
    return vd::synth_bsc_launch(h->geom, h->labels_dev, nframes, first_frame, seed, crossover_ppm, payload_dev, symbols_dev,
                                static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
