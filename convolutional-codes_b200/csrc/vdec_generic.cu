// vdec_generic.cu — table-driven Viterbi kernels for ANY code with K <= 15.
//
// Serves every trellis the specialised register kernels (vdec_regs.cuh) do not cover: K >= 8
// (256 .. 16384 states, BASELINE config 4) and arbitrary polynomials.  Same arithmetic and tie
// rules as the reference (AWGN-channel/viterbi-decoder.c:108-141, 71-90;
// binary-symmetric-channel/viterbi-decoder.c:114-150, 75-96).  Two launches per batch:
//
//  1. vit_generic_forward — ONE CTA PER FRAME.
//     * path metrics double-buffered in shared memory (2 x S x 4 B; K=15: 128 KB);
//     * per-butterfly labels (4 bytes: l(2m,0), l(2m+1,0), l(2m,1), l(2m+1,1)), generated on the
//       host from the caller's polynomials: in registers when every thread owns one butterfly
//       (S/2 <= block size), else staged once per CTA in shared memory (K=15: 32 KB);
//     * the step's M branch metrics sit in a double-buffered shared array (written one step ahead),
//       so `bm[label]` is one broadcast LDS instead of a select chain;
//     * thread m handles butterflies m, m+NT, ...: reads (pm[2m], pm[2m+1]) as one 8-byte access,
//       writes pm'[m] and pm'[m+S/2]; decisions are gathered with __ballot_sync into 32-bit
//       survivor words, one bit per state per step, stored to HBM ([frame][step][S/32]);
//     * final state: lowest-index argmin (block reduction on (metric, index)) -> final_state[frame].
//
//  2. vit_generic_traceback — ONE THREAD PER FRAME.
//     The predecessor recurrence cur' = ((cur << 1) & (S-1)) | d shifts new decision bits in at the
//     LSB, so the survivor WORD needed k steps later, (cur_k >> 5), is known from the current state
//     for k <= 5.  Each thread therefore keeps five dependent-free loads in flight and only the bit
//     select is on the serial chain; thousands of frames hide the remaining latency.  Inside a CTA the
//     same walk costs ~1 us per step and dominated the first version of this file by 20x.
#include <stdio.h>
#include <stdlib.h>

#include "vdec_internal.h"

namespace vd {

namespace {

struct SoftOps {
    using T = float;
    __device__ static T unreachable() { return __int_as_float(0x7f800000); }
    __device__ static T cand(T pm, T bm) { return __fadd_rn(pm, bm); }                        // :121
};
struct HardOps {
    using T = uint32_t;
    __device__ static T unreachable() { return 0xFF00u; }
    __device__ static T cand(T pm, T bm) { return min(pm + bm, 0xFF00u); }                     // :127-130
};

struct QuantOps {             // VDEC_SOFT_U8: exact integer sums, no clamp (the float decoder has none)
    using T = uint32_t;
    __device__ static T unreachable() { return 0x3FFFFFFFu; }
    __device__ static T cand(T pm, T bm) { return pm + bm; }
};

// surv layout: [frame][step][nw] words; tail: final_state[nframes] words
// ONE_EACH: S/2 <= blockDim, every thread owns (at most) one butterfly for the whole frame.
// IN: 0 = float metrics, 1 = hard symbols (Hamming), 2 = uint8 metrics
template <class OPS, int IN, bool ONE_EACH>
__global__ void __launch_bounds__(1024)
vit_generic_forward(LaunchGeom g, const uint32_t* __restrict__ bfly_labels, const void* __restrict__ in_v,
                    int32_t* __restrict__ path_metric, uint32_t* __restrict__ surv, uint32_t* __restrict__ final_state) {
    using T = typename OPS::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* pm0 = reinterpret_cast<T*>(smem_raw);
    T* pm1 = pm0 + g.S;
    uint32_t* lab = reinterpret_cast<uint32_t*>(pm1 + g.S);     // [S/2], only when a thread owns > 1 butterfly
    __shared__ T bm_s[2][8];
    __shared__ T red_val[32];
    __shared__ uint32_t red_idx[32];

    const uint32_t f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    const uint32_t half = g.S >> 1, nw = g.nw, T_ = g.T, M = g.M;
    uint32_t* sv = surv + (size_t)f * T_ * nw;

    for (uint32_t s = tid; s < g.S; s += nt) pm0[s] = (s == 0 || f >= g.init_all_from) ? T(0) : OPS::unreachable();          // :215-218
    if constexpr (!ONE_EACH) for (uint32_t m = tid; m < half; m += nt) lab[m] = __ldg(bfly_labels + m);

    // threads 0..M-1 publish the branch metrics one step ahead of their use
    const float* in_f = static_cast<const float*>(in_v) + (size_t)f * g.stride_steps * M + tid;
    const uint8_t* in_b = static_cast<const uint8_t*>(in_v) + (size_t)f * g.stride_steps;
    const uint8_t* in_q = static_cast<const uint8_t*>(in_v) + (size_t)f * g.stride_steps * M + tid;
    const bool stager = tid < M;
    auto stage_bm = [&](uint32_t t) {
        if (stager && t < T_) {
            if constexpr (IN == 1) bm_s[t & 1][tid] = __popc(tid ^ (__ldg(in_b + t) & (M - 1)));   // :69-72, :235-237
            else if constexpr (IN == 2) bm_s[t & 1][tid] = __ldg(in_q + (size_t)t * M);
            else bm_s[t & 1][tid] = __ldg(in_f + (size_t)t * M);
        }
    };
    stage_bm(0);
    __syncthreads();

    T* cur = pm0; T* nxt = pm1;
    if constexpr (ONE_EACH) {
        const bool active = tid < half;
        const uint32_t m = active ? tid : 0;
        const uint32_t lb = __ldg(bfly_labels + m);
        const uint32_t l00 = lb & 0xFF, l01 = (lb >> 8) & 0xFF, l10 = (lb >> 16) & 0xFF, l11 = lb >> 24;
        // survivor word slots this warp's lane 0 writes
        const bool writer = active && lane == 0;
        uint32_t* w0p = sv + (half >= 32 ? (m >> 5) : 0);
        uint32_t* w1p = sv + ((m + half) >> 5);
        for (uint32_t t = 0; t < T_; ++t) {
            stage_bm(t + 1);
            const T* bm = bm_s[t & 1];
            const T pe = cur[2 * m], po = cur[2 * m + 1];
            const T c00 = OPS::cand(pe, bm[l00]), c01 = OPS::cand(po, bm[l01]);
            const T c10 = OPS::cand(pe, bm[l10]), c11 = OPS::cand(po, bm[l11]);
            const bool d0 = active && (c01 < c00), d1 = active && (c11 < c10);                   // strict: tie -> even (:122)
            if (active) { nxt[m] = d0 ? c01 : c00; nxt[m + half] = d1 ? c11 : c10; }
            const uint32_t b0 = __ballot_sync(0xffffffffu, d0), b1 = __ballot_sync(0xffffffffu, d1);
            if (writer) {
                if (half >= 32) { *w0p = b0; *w1p = b1; }
                else *w0p = b0 | (b1 << half);                  // S <= 32: one word per step
            }
            w0p += nw; w1p += nw;
            __syncthreads();
            T* tmp = cur; cur = nxt; nxt = tmp;
        }
    } else {
        // half is a multiple of nt (both powers of two, half > nt)
        for (uint32_t t = 0; t < T_; ++t) {
            stage_bm(t + 1);
            const T* bm = bm_s[t & 1];
            uint32_t* row = sv + (size_t)t * nw;
#pragma unroll 4
            for (uint32_t m = tid; m < half; m += nt) {
                const uint32_t lb = lab[m];
                const T pe = cur[2 * m], po = cur[2 * m + 1];
                const T c00 = OPS::cand(pe, bm[lb & 0xFF]), c01 = OPS::cand(po, bm[(lb >> 8) & 0xFF]);
                const T c10 = OPS::cand(pe, bm[(lb >> 16) & 0xFF]), c11 = OPS::cand(po, bm[lb >> 24]);
                const bool d0 = c01 < c00, d1 = c11 < c10;
                nxt[m] = d0 ? c01 : c00;
                nxt[m + half] = d1 ? c11 : c10;
                const uint32_t b0 = __ballot_sync(0xffffffffu, d0), b1 = __ballot_sync(0xffffffffu, d1);
                if (lane == 0) { row[m >> 5] = b0; row[(m + half) >> 5] = b1; }
            }
            __syncthreads();
            T* tmp = cur; cur = nxt; nxt = tmp;
        }
    }

    // final state: lowest-index argmin with strict '<' from the sentinel (:72-79)
    T best = OPS::unreachable(); uint32_t idx = 0xFFFFFFFFu;
    for (uint32_t s = tid; s < g.S; s += nt) {
        const T v = cur[s];
        if (v < best) { best = v; idx = s; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const T ov = __shfl_down_sync(0xffffffffu, best, off);
        const uint32_t oi = __shfl_down_sync(0xffffffffu, idx, off);
        if (ov < best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if (lane == 0) { red_val[tid >> 5] = best; red_idx[tid >> 5] = idx; }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t w = 1; w < (nt + 31) / 32; ++w)
            if (red_val[w] < best || (red_val[w] == best && red_idx[w] < idx)) { best = red_val[w]; idx = red_idx[w]; }
        if (idx == 0xFFFFFFFFu) idx = 0;                       // nothing below the sentinel -> state 0
        final_state[f] = idx;
        if (IN == 1 && path_metric != nullptr) path_metric[f] = (int32_t)best;                       // :94,:143
    }
}

// ---- forward kernel for S >= 64: BPT butterflies per thread, everything loop-invariant in registers ----
// nt = (S/2)/BPT threads (a multiple of 32), thread `tid` owns butterflies m = i*nt + tid, i < BPT.  Labels
// are unpacked once into registers, the two path-metric buffers and the two branch-metric buffers are
// addressed by constants (the step loop is unrolled by two), and a warp's ballots of butterfly row i land
// in survivor words (i*nt/32 + warp) and (S/64 + i*nt/32 + warp): natural bit order, bit s&31 of word s>>5.
template <class OPS, int IN, int BPT>
__global__ void __launch_bounds__(1024)
vit_generic_forward_v2(LaunchGeom g, const uint32_t* __restrict__ bfly_labels, const void* __restrict__ in_v,
                       int32_t* __restrict__ path_metric, uint32_t* __restrict__ surv, uint32_t* __restrict__ final_state) {
    using T = typename OPS::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* pm0 = reinterpret_cast<T*>(smem_raw);
    T* pm1 = pm0 + g.S;
    __shared__ T bm_s[2][8];
    __shared__ T red_val[32];
    __shared__ uint32_t red_idx[32];

    const uint32_t f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
    const uint32_t half = g.S >> 1, nw = g.nw, T_ = g.T, M = g.M;

    uint32_t l00[BPT], l01[BPT], l10[BPT], l11[BPT];
#pragma unroll
    for (int i = 0; i < BPT; ++i) {
        const uint32_t lb = __ldg(bfly_labels + i * nt + tid);
        l00[i] = lb & 0xFF; l01[i] = (lb >> 8) & 0xFF; l10[i] = (lb >> 16) & 0xFF; l11[i] = lb >> 24;
    }
    for (uint32_t s = tid; s < g.S; s += nt) pm0[s] = (s == 0 || f >= g.init_all_from) ? T(0) : OPS::unreachable();          // :215-218

    const bool stager = tid < M;
    const float* in_f = static_cast<const float*>(in_v) + (size_t)f * g.stride_steps * M + tid;
    const uint8_t* in_b = static_cast<const uint8_t*>(in_v) + (size_t)f * g.stride_steps;
    const uint8_t* in_q = static_cast<const uint8_t*>(in_v) + (size_t)f * g.stride_steps * M + tid;
    auto stage_bm = [&](uint32_t t, T* dst) {                   // step t's metrics, one step ahead of their use
        if (stager) {
            const uint32_t tc = t < T_ ? t : T_ - 1;
            if constexpr (IN == 1) dst[tid] = __popc(tid ^ (__ldg(in_b + tc) & (M - 1)));        // :69-72, :235-237
            else if constexpr (IN == 2) dst[tid] = __ldg(in_q + (size_t)tc * M);
            else dst[tid] = __ldg(in_f + (size_t)tc * M);
        }
    };
    uint32_t* row = surv + (size_t)f * T_ * nw + warp;          // lane 0 of each warp writes its words
    const uint32_t wstride = nt >> 5, hi_off = half >> 5;
    auto step = [&](const T* __restrict__ cur, T* __restrict__ nxt, const T* __restrict__ bm) {
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const uint32_t m = i * nt + tid;
            const T pe = cur[2 * m], po = cur[2 * m + 1];
            const T c00 = OPS::cand(pe, bm[l00[i]]), c01 = OPS::cand(po, bm[l01[i]]);
            const T c10 = OPS::cand(pe, bm[l10[i]]), c11 = OPS::cand(po, bm[l11[i]]);
            const bool d0 = c01 < c00, d1 = c11 < c10;                                            // strict: tie -> even (:122)
            nxt[m] = d0 ? c01 : c00;
            nxt[m + half] = d1 ? c11 : c10;
            const uint32_t b0 = __ballot_sync(0xffffffffu, d0), b1 = __ballot_sync(0xffffffffu, d1);
            if (lane == 0) { row[i * wstride] = b0; row[hi_off + i * wstride] = b1; }
        }
        row += nw;
    };
    stage_bm(0, bm_s[0]);
    __syncthreads();
    uint32_t t = 0;
    for (; t + 2 <= T_; t += 2) {
        stage_bm(t + 1, bm_s[1]);
        step(pm0, pm1, bm_s[0]);
        __syncthreads();
        stage_bm(t + 2, bm_s[0]);
        step(pm1, pm0, bm_s[1]);
        __syncthreads();
    }
    const T* fin = pm0;
    if (t < T_) { step(pm0, pm1, bm_s[0]); __syncthreads(); fin = pm1; }

    // final state: lowest-index argmin with strict '<' from the sentinel (:72-79)
    T best = OPS::unreachable(); uint32_t idx = 0xFFFFFFFFu;
    for (uint32_t s = tid; s < g.S; s += nt) {
        const T v = fin[s];
        if (v < best) { best = v; idx = s; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const T ov = __shfl_down_sync(0xffffffffu, best, off);
        const uint32_t oi = __shfl_down_sync(0xffffffffu, idx, off);
        if (ov < best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if (lane == 0) { red_val[warp] = best; red_idx[warp] = idx; }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t w = 1; w < (nt + 31) / 32; ++w)
            if (red_val[w] < best || (red_val[w] == best && red_idx[w] < idx)) { best = red_val[w]; idx = red_idx[w]; }
        if (idx == 0xFFFFFFFFu) idx = 0;
        final_state[f] = idx;
        if (IN == 1 && path_metric != nullptr) path_metric[f] = (int32_t)best;                    // :94,:143
    }
}

constexpr int kDepth = 5;      // survivor words in flight per thread (see file header)

__global__ void __launch_bounds__(128)
vit_generic_traceback(LaunchGeom g, const uint32_t* __restrict__ surv, const uint32_t* __restrict__ final_state,
                      uint8_t* __restrict__ out) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= g.nframes) return;
    const uint32_t* sv = surv + (size_t)f * g.T * g.nw;
    uint8_t* o = out + (size_t)f * g.out_bytes;
    uint32_t c = final_state[f];
    const uint32_t smask = g.S - 1, msb = g.K - 2;
    // pipe[k] = word needed at step (t - k), whose index is known from the state at step t
    uint32_t pipe[kDepth];
    int t = (int)g.T - 1;
#pragma unroll
    for (int k = 0; k < kDepth; ++k) {
        const int tk = t - k;
        pipe[k] = tk >= 0 ? __ldcg(sv + (size_t)tk * g.nw + (((c << k) & smask) >> 5)) : 0u;
    }
    uint32_t acc = 0;
    while (t >= 0) {
        // consume kDepth steps from the pipe, refilling each slot with the word kDepth steps further back
#pragma unroll
        for (int k = 0; k < kDepth; ++k) {
            if (t >= 0) {
                const uint32_t word = pipe[k];
                const uint32_t d = (word >> (c & 31)) & 1u;
                acc |= (c >> msb) << (7 - (t & 7));                                                 // :88
                c = ((c << 1) & smask) | d;                                                          // :86
                if ((t & 7) == 0) { o[t >> 3] = (uint8_t)acc; acc = 0; }
                --t;
                // state at step t (just computed) fixes the word index kDepth-1 steps further back
                const int tn = t - (kDepth - 1);
                pipe[k] = tn >= 0 ? __ldcg(sv + (size_t)tn * g.nw + (((c << (kDepth - 1)) & smask) >> 5)) : 0u;
            }
        }
    }
}

}  // namespace

size_t generic_surv_bytes(const LaunchGeom& g) {
    return ((size_t)g.nframes * g.T * g.nw + g.nframes) * sizeof(uint32_t);       // survivors + final states
}
size_t generic_smem_bytes(const LaunchGeom& g) { return (size_t)g.S * 4 * 2 + (size_t)(g.S / 2) * 4; }

template <class OPS, int IN, int BPT>
static int launch_v2(const LaunchGeom& g, const uint32_t* labels, const void* in, int32_t* path_metric, uint32_t* surv,
                     uint32_t* final_state, cudaStream_t st) {
    const size_t smem = (size_t)g.S * 4 * 2;
    auto k = vit_generic_forward_v2<OPS, IN, BPT>;
    if (smem > 48 * 1024) VDEC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<g.nframes, (g.S / 2) / BPT, smem, st>>>(g, labels, in, path_metric, surv, final_state);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

template <class OPS, int IN>
static int generic_forward_launch(const LaunchGeom& g, const uint32_t* labels, const void* in, int32_t* path_metric,
                                  uint32_t* surv, uint32_t* final_state, cudaStream_t st) {
    const uint32_t half = g.S / 2;
    if (half < 64) {                 // tiny trellises: one (partly idle) warp or two per frame, simple path
        const size_t smem = generic_smem_bytes(g);
        auto k = vit_generic_forward<OPS, IN, true>;
        k<<<g.nframes, half < 32 ? 32 : half, smem, st>>>(g, labels, in, path_metric, surv, final_state);
        VDEC_CUDA(cudaGetLastError());
        return 0;
    }
    // butterflies per thread: tuned on B200 (tools/bench_configs.py); VDEC_GENERIC_BPT overrides
    uint32_t bpt = half <= 128 ? 2 : half <= 1024 ? 4 : 8;
    if (const char* e = getenv("VDEC_GENERIC_BPT")) bpt = (uint32_t)atoi(e);
    while (bpt > 1 && half / bpt < 32) bpt >>= 1;
    while (half / bpt > 1024) bpt <<= 1;
    switch (bpt) {
        case 1: return launch_v2<OPS, IN, 1>(g, labels, in, path_metric, surv, final_state, st);
        case 2: return launch_v2<OPS, IN, 2>(g, labels, in, path_metric, surv, final_state, st);
        case 4: return launch_v2<OPS, IN, 4>(g, labels, in, path_metric, surv, final_state, st);
        case 8: return launch_v2<OPS, IN, 8>(g, labels, in, path_metric, surv, final_state, st);
        default: set_error("unsupported butterflies-per-thread %u", bpt); return -1;
    }
}

int generic_launch(const LaunchGeom& g, uint32_t metric, const uint32_t* bfly_labels_dev, const void* in,
                   uint8_t* out, int32_t* path_metric, uint32_t* surv, cudaStream_t st) {
    uint32_t* final_state = surv + (size_t)g.nframes * g.T * g.nw;
    const int rc = metric == VDEC_HARD_U16  ? generic_forward_launch<HardOps, 1>(g, bfly_labels_dev, in, path_metric, surv, final_state, st)
                   : metric == VDEC_SOFT_U8 ? generic_forward_launch<QuantOps, 2>(g, bfly_labels_dev, in, path_metric, surv, final_state, st)
                                            : generic_forward_launch<SoftOps, 0>(g, bfly_labels_dev, in, path_metric, surv, final_state, st);
    if (rc != 0) return -1;
    vit_generic_traceback<<<(g.nframes + 127) / 128, 128, 0, st>>>(g, surv, final_state, out);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vd
