// vdec_sweep.cu — the drivers' operating-point loop, sharded over the GPUs of one node, behind ONE C call.
//
// Replaces the loops of AWGN-channel/main.c:193-234 (Eb/N0 grid, per point: frames through encoder -> mapper ->
// channel -> demapper -> decoder, bit errors accumulated in data_callback_decoded :129-141, BER printed :233) and of
// binary-symmetric-channel/main.c:147-176 (17 crossover rates) for callers that stay in C: single process, one
// stream per device, frames of a point split into contiguous ranges of the global frame index (the same split as
// link.py:shard_range, so the Python and C drivers count the same errors), and ONE collective per batch — an
// ncclAllReduce(sum, uint64) over the three counters {bits, bit errors, frame errors}.
// NCCL is opened with dlopen on first multi-GPU use: the library keeps no link-time dependency on it, and a process
// that already carries an NCCL (PyTorch) shares that copy.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "vdec_internal.h"

namespace vd {
namespace {

struct Nccl {
    void* so = nullptr;
    bool tried = false;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl g_nccl;

bool load_nccl() {
    if (g_nccl.tried) return g_nccl.so != nullptr;
    g_nccl.tried = true;
    const char* names[] = {getenv("VDEC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        g_nccl.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.so) break;
    }
    if (!g_nccl.so) { set_error("libnccl not found (set VDEC_NCCL_LIB): %s", dlerror()); return false; }
    bool ok = true;
#define VDEC_SYM(f) ok = ok && ((*(void**)(&g_nccl.f) = dlsym(g_nccl.so, "nccl" #f)) != nullptr)
    VDEC_SYM(CommInitAll); VDEC_SYM(CommDestroy); VDEC_SYM(AllReduce); VDEC_SYM(GroupStart); VDEC_SYM(GroupEnd); VDEC_SYM(GetErrorString);
#undef VDEC_SYM
    if (!ok) { set_error("libnccl lacks a required symbol"); dlclose(g_nccl.so); g_nccl.so = nullptr; }
    return ok;
}

#define VDEC_NCCL(call)                                                                                  \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) { set_error("%s failed: %s", #call, g_nccl.GetErrorString(r__)); return -1; } \
    } while (0)

// contiguous [first, first + count) of `total` frames for shard `rank` of `world` (link.py:shard_range)
void shard_range(uint64_t total, int rank, int world, uint64_t* first, uint64_t* count) {
    const uint64_t base = total / world, extra = total % world;
    *first = rank * base + ((uint64_t)rank < extra ? (uint64_t)rank : extra);
    *count = base + ((uint64_t)rank < extra ? 1 : 0);
}

struct Dev {
    int id = -1;
    vdec_t* dec = nullptr;
    cudaStream_t st = nullptr;
    uint8_t* payload = nullptr;
    void* chan = nullptr;          // I/Q (AWGN) or symbol bytes (BSC)
    float* met = nullptr;          // hard-demapper output
    uint8_t* out = nullptr;
    uint64_t* counters = nullptr;  // [3], this batch
};

// communicators are created once per device list and kept for the life of the process (ncclCommInitAll costs ~1 s)
struct CommSet { std::vector<int> ids; std::vector<ncclComm_t> comm; };
std::vector<CommSet> g_comms;
std::mutex g_sweep_mu;                 // one sweep at a time per process: the communicators and their streams are shared

// the caller's current device is put back when the sweep returns (it runs cudaSetDevice for every device it uses)
struct DeviceRestore {
    int dev = -1;
    DeviceRestore() { if (cudaGetDevice(&dev) != cudaSuccess) dev = -1; }
    ~DeviceRestore() { if (dev >= 0) cudaSetDevice(dev); }
};

struct Sweep {
    std::vector<Dev> dev;
    ncclComm_t* comm = nullptr;
    ~Sweep() {
        for (Dev& d : dev) {
            if (d.id < 0) continue;
            cudaSetDevice(d.id);
            if (d.st) cudaStreamSynchronize(d.st);
            cudaFree(d.payload); cudaFree(d.chan); cudaFree(d.met); cudaFree(d.out); cudaFree(d.counters);
            if (d.dec) vdec_destroy(d.dec);
            if (d.st) cudaStreamDestroy(d.st);
        }
    }
};

// resume file: one line per finished batch, "point frames_done bits bit_errors frame_errors"
void resume_load(const char* path, uint32_t npoints, std::vector<uint64_t>* done, uint64_t* counters) {
    FILE* f = fopen(path, "r");
    if (!f) return;
    unsigned long long i, fr, b, be, fe;
    while (fscanf(f, "%llu %llu %llu %llu %llu", &i, &fr, &b, &be, &fe) == 5)
        if (i < npoints && fr >= (*done)[i]) { (*done)[i] = fr; counters[3 * i] = b; counters[3 * i + 1] = be; counters[3 * i + 2] = fe; }
    fclose(f);
}

}  // namespace
}  // namespace vd

extern "C" int vdec_sweep(const struct vdec_config* cfg, const struct vdec_sweep_config* sw, const double* points, uint32_t npoints,
                          const int* devices, int ndev, uint64_t* counters_out) {
    using namespace vd;
    if (!cfg || !sw || !points || !counters_out || npoints == 0) { set_error("vdec_sweep: null argument"); return -1; }
    if (sw->channel > VDEC_CHANNEL_BSC) { set_error("vdec_sweep: unknown channel %u", sw->channel); return -1; }
    if (sw->frames_per_batch == 0) { set_error("vdec_sweep: frames_per_batch is 0"); return -1; }
    int present = 0;
    if (cudaGetDeviceCount(&present) != cudaSuccess || present == 0) { set_error("no CUDA device: this library has no CPU path"); return -1; }
    if (ndev <= 0) ndev = present;                                  // all devices of the node
    if (ndev > present) { set_error("vdec_sweep: %d devices asked for, %d present", ndev, present); return -1; }
    const bool bsc = sw->channel == VDEC_CHANNEL_BSC, hard_demap = sw->channel == VDEC_CHANNEL_AWGN_HARD;
    if ((cfg->metric == VDEC_HARD_U16) != bsc) { set_error("vdec_sweep: the BSC channel needs VDEC_HARD_U16, the AWGN channels VDEC_SOFT_F32"); return -1; }
    if (!bsc && cfg->metric != VDEC_SOFT_F32) { set_error("vdec_sweep: the AWGN channels need VDEC_SOFT_F32"); return -1; }

    std::lock_guard<std::mutex> lock(g_sweep_mu);
    DeviceRestore restore;                                          // destroyed after `s`, whose destructor switches devices
    Sweep s;
    s.dev.resize(ndev);
    // frames per launch and device: four full waves of the register kernels (148 SMs x 4 CTAs x 128 frames)
    const uint32_t chunk = sw->chunk_frames ? sw->chunk_frames : 4u * 148u * 4u * 128u;
    for (int r = 0; r < ndev; ++r) {
        Dev& d = s.dev[r];
        vdec_config c = *cfg;
        c.device = devices ? devices[r] : r;
        VDEC_CUDA(cudaSetDevice(c.device));
        d.id = c.device;
        if (vdec_create(&d.dec, &c) != 0) return -1;
        const LaunchGeom& g = d.dec->geom;
        VDEC_CUDA(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
        VDEC_CUDA(cudaMalloc(&d.payload, (size_t)chunk * ((g.L + 7) / 8)));
        VDEC_CUDA(cudaMalloc(&d.chan, (size_t)chunk * g.T * (bsc ? 1 : 2 * sizeof(float))));
        if (hard_demap) VDEC_CUDA(cudaMalloc(&d.met, (size_t)chunk * g.T * g.M * sizeof(float)));
        VDEC_CUDA(cudaMalloc(&d.out, (size_t)chunk * g.out_bytes));
        VDEC_CUDA(cudaMalloc(&d.counters, 3 * sizeof(uint64_t)));
    }
    if (ndev > 1) {
        if (!load_nccl()) return -1;
        std::vector<int> ids(ndev);
        for (int r = 0; r < ndev; ++r) ids[r] = s.dev[r].id;
        for (CommSet& c : g_comms) if (c.ids == ids) s.comm = c.comm.data();
        if (!s.comm) {
            CommSet fresh{ids, std::vector<ncclComm_t>(ndev, nullptr)};
            VDEC_NCCL(g_nccl.CommInitAll(fresh.comm.data(), ndev, ids.data()));     // nothing is cached when this fails
            g_comms.push_back(std::move(fresh));
            s.comm = g_comms.back().comm.data();
        }
    }

    std::vector<uint64_t> done(npoints, 0);
    memset(counters_out, 0, sizeof(uint64_t) * 3 * npoints);
    if (sw->resume_path && *sw->resume_path) resume_load(sw->resume_path, npoints, &done, counters_out);

    for (uint32_t i = 0; i < npoints; ++i) {
        uint64_t* tot = counters_out + 3 * i;
        const uint64_t seed = sw->seed + 7919ull * i;                       // link.py:sweep_awgn
        const float sigma = bsc ? 0.f : (float)(0.70710678118654752440 * pow(10.0, -points[i] / 20.0));   // main.c:153-161
        const uint32_t ppm = bsc ? (uint32_t)(points[i] + 0.5) : 0u;
        for (;;) {
            const bool enough = done[i] > 0 && (sw->min_frame_errors == 0 || tot[2] >= sw->min_frame_errors ||
                                                (sw->max_frames && done[i] >= sw->max_frames));
            if (enough) break;
            // one batch: every device takes its range of [done, done + frames_per_batch)
            for (int r = 0; r < ndev; ++r) {
                VDEC_CUDA(cudaSetDevice(s.dev[r].id));
                VDEC_CUDA(cudaMemsetAsync(s.dev[r].counters, 0, 3 * sizeof(uint64_t), s.dev[r].st));
            }
            uint64_t most = 0;
            std::vector<uint64_t> first(ndev), count(ndev);
            for (int r = 0; r < ndev; ++r) { shard_range(sw->frames_per_batch, r, ndev, &first[r], &count[r]); if (count[r] > most) most = count[r]; }
            for (uint64_t off = 0; off < most; off += chunk) {                // devices interleaved: all of them stay busy
                for (int r = 0; r < ndev; ++r) {
                    if (off >= count[r]) continue;
                    Dev& d = s.dev[r];
                    const uint64_t n = count[r] - off < chunk ? count[r] - off : chunk;
                    const uint64_t f0 = done[i] + first[r] + off;
                    const LaunchGeom& g = d.dec->geom;
                    if (bsc) {
                        if (vdec_synth_bsc_device(d.dec, n, f0, seed, ppm, d.payload, static_cast<uint8_t*>(d.chan), d.st) != 0) return -1;
                        if (vdec_decode_device(d.dec, d.chan, n, d.out, nullptr, d.st) != 0) return -1;
                    } else {
                        float* iq = static_cast<float*>(d.chan);
                        if (vdec_synth_awgn_device(d.dec, n, f0, seed, sigma, d.payload, nullptr, iq, d.st) != 0) return -1;
                        if (hard_demap) {
                            if (vdec_hard_demap_device(g.n, iq, n * g.T, d.met, d.st) != 0) return -1;
                            if (vdec_decode_device(d.dec, d.met, n, d.out, nullptr, d.st) != 0) return -1;
                        } else if (vdec_decode_iq_device(d.dec, iq, n, d.out, d.st) != 0) return -1;
                    }
                    if (vdec_count_errors_device(d.dec, d.out, d.payload, n, d.counters, d.st) != 0) return -1;
                }
            }
            if (ndev > 1) {                                                   // the path's one collective
                VDEC_NCCL(g_nccl.GroupStart());
                for (int r = 0; r < ndev; ++r)
                    VDEC_NCCL(g_nccl.AllReduce(s.dev[r].counters, s.dev[r].counters, 3, ncclUint64, ncclSum, s.comm[r], s.dev[r].st));
                VDEC_NCCL(g_nccl.GroupEnd());
            }
            uint64_t batch[3];
            VDEC_CUDA(cudaSetDevice(s.dev[0].id));
            VDEC_CUDA(cudaMemcpyAsync(batch, s.dev[0].counters, sizeof(batch), cudaMemcpyDeviceToHost, s.dev[0].st));
            for (int r = 0; r < ndev; ++r) { VDEC_CUDA(cudaSetDevice(s.dev[r].id)); VDEC_CUDA(cudaStreamSynchronize(s.dev[r].st)); }
            for (int k = 0; k < 3; ++k) tot[k] += batch[k];
            done[i] += sw->frames_per_batch;
            if (sw->resume_path && *sw->resume_path) {
                if (FILE* f = fopen(sw->resume_path, "a")) {
                    fprintf(f, "%u %llu %llu %llu %llu\n", i, (unsigned long long)done[i], (unsigned long long)tot[0],
                            (unsigned long long)tot[1], (unsigned long long)tot[2]);
                    fclose(f);
                }
            }
        }
    }
    return 0;
}

extern "C" int vdec_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
