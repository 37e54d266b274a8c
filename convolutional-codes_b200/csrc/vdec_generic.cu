// vdec_generic.cu — table-driven Viterbi kernel for ANY code with K <= 15: ONE CTA PER FRAME.
//
// Serves every trellis the specialised register kernels (vdec_regs.cuh) do not cover: K >= 8
// (256 .. 16384 states, BASELINE config 4) and arbitrary polynomials.  Same arithmetic and tie
// rules as the reference (AWGN-channel/viterbi-decoder.c:108-141, 71-90;
// binary-symmetric-channel/viterbi-decoder.c:114-150, 75-96):
//   * path metrics double-buffered in shared memory (2 x S x 4 B; K=15: 128 KB),
//   * per-butterfly labels (4 bytes: l(2m,0), l(2m+1,0), l(2m,1), l(2m+1,1)) staged once per CTA
//     in shared memory (K=15: 32 KB), generated on the host from the caller's polynomials,
//   * thread m handles butterflies m, m+NT, ...: reads (pm[2m], pm[2m+1]) as one 8-byte access,
//     writes pm'[m] and pm'[m+S/2]; decisions are gathered with __ballot_sync into 32-bit survivor
//     words, one bit per state per step, and stored to an HBM scratch ([frame][step][S/32]),
//   * final state: lowest-index argmin (block reduction on (metric, index)),
//   * traceback by warp 0 lane 0 over the scratch (latency-bound; a few % of the forward pass
//     for the large-K configurations this kernel is for).
#include <stdio.h>

#include "vdec_internal.h"

namespace vd {

namespace {

constexpr int kGenThreads = 256;

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

template <class OPS, bool HARD>
__global__ void __launch_bounds__(kGenThreads)
vit_generic_kernel(LaunchGeom g, const uint32_t* __restrict__ bfly_labels, const void* __restrict__ in_v,
                   uint8_t* __restrict__ out, int32_t* __restrict__ path_metric, uint32_t* __restrict__ surv) {
    using T = typename OPS::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* pm0 = reinterpret_cast<T*>(smem_raw);
    T* pm1 = pm0 + g.S;
    uint32_t* lab = reinterpret_cast<uint32_t*>(pm1 + g.S);     // [S/2]
    __shared__ T red_val[kGenThreads / 32];
    __shared__ uint32_t red_idx[kGenThreads / 32];
    __shared__ uint32_t s_cur;

    const uint32_t f = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const uint32_t half = g.S >> 1;
    uint32_t* sv = surv + (size_t)f * g.T * g.nw;

    for (uint32_t m = tid; m < half; m += kGenThreads) lab[m] = __ldg(bfly_labels + m);
    for (uint32_t s = tid; s < g.S; s += kGenThreads) pm0[s] = (s == 0) ? T(0) : OPS::unreachable();   // :215-218
    __syncthreads();

    T* cur = pm0; T* nxt = pm1;
    const float* in_f = static_cast<const float*>(in_v) + (size_t)f * g.T * g.M;
    const uint8_t* in_b = static_cast<const uint8_t*>(in_v) + (size_t)f * g.T;
    const uint32_t iters = (half + kGenThreads - 1) / kGenThreads;

    for (uint32_t t = 0; t < g.T; ++t) {
        T bm[8];
        if constexpr (HARD) {
            const uint32_t r = __ldg(in_b + t) & (g.M - 1);                                      // :235-237
#pragma unroll
            for (uint32_t e = 0; e < 8; ++e) bm[e] = __popc(e ^ r);                              // :69-72
        } else {
            const float* p = in_f + (size_t)t * g.M;
#pragma unroll
            for (uint32_t e = 0; e < 8; ++e) bm[e] = (e < g.M) ? __ldg(p + e) : 0.f;
        }
        for (uint32_t it = 0; it < iters; ++it) {
            const uint32_t m = it * kGenThreads + tid;
            bool d0 = false, d1 = false;
            if (m < half) {
                const uint32_t lb = lab[m];
                const T pe = cur[2 * m], po = cur[2 * m + 1];
                // register array indexed dynamically -> select chain (M <= 8)
                auto pick = [&](uint32_t e) -> T {
                    T v = bm[0];
#pragma unroll
                    for (uint32_t k = 1; k < 8; ++k) v = (e == k) ? bm[k] : v;
                    return v;
                };
                const T c00 = OPS::cand(pe, pick(lb & 0xFF)), c01 = OPS::cand(po, pick((lb >> 8) & 0xFF));
                const T c10 = OPS::cand(pe, pick((lb >> 16) & 0xFF)), c11 = OPS::cand(po, pick(lb >> 24));
                d0 = c01 < c00; d1 = c11 < c10;                                                   // strict: tie -> even (:122)
                nxt[m] = d0 ? c01 : c00;
                nxt[m + half] = d1 ? c11 : c10;
            }
            const uint32_t w0 = __ballot_sync(0xffffffffu, d0), w1 = __ballot_sync(0xffffffffu, d1);
            if (lane == 0 && (m < half)) {
                if (half >= 32) { sv[(size_t)t * g.nw + (m >> 5)] = w0; sv[(size_t)t * g.nw + ((m + half) >> 5)] = w1; }
                else sv[(size_t)t * g.nw] = w0 | (w1 << half);          // S <= 32: one word per step
            }
        }
        __syncthreads();
        T* tmp = cur; cur = nxt; nxt = tmp;
    }

    // final state: lowest-index argmin with strict '<' from the sentinel (:72-79)
    T best = OPS::unreachable(); uint32_t idx = 0xFFFFFFFFu;
    for (uint32_t s = tid; s < g.S; s += kGenThreads) {
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
        for (int w = 1; w < kGenThreads / 32; ++w)
            if (red_val[w] < best || (red_val[w] == best && red_idx[w] < idx)) { best = red_val[w]; idx = red_idx[w]; }
        if (idx == 0xFFFFFFFFu) idx = 0;                       // nothing below the sentinel -> state 0
        s_cur = idx;
        if (HARD && path_metric != nullptr) path_metric[f] = (int32_t)best;                       // :94,:143
    }
    __syncthreads();

    // traceback (:81-89)
    if (tid == 0) {
        uint32_t c = s_cur;
        uint8_t* o = out + (size_t)f * g.out_bytes;
        uint32_t acc = 0;
        for (int t = (int)g.T - 1; t >= 0; --t) {
            const uint32_t word = __ldcg(sv + (size_t)t * g.nw + (c >> 5));
            const uint32_t d = (word >> (c & 31)) & 1u;
            acc |= (c >> (g.K - 2)) << (7 - (t & 7));
            c = ((c << 1) & (g.S - 1)) | d;
            if ((t & 7) == 0) { o[t >> 3] = (uint8_t)acc; acc = 0; }
        }
    }
}

}  // namespace

size_t generic_surv_bytes(const LaunchGeom& g) { return (size_t)g.nframes * g.T * g.nw * sizeof(uint32_t); }
size_t generic_smem_bytes(const LaunchGeom& g) { return (size_t)g.S * 4 * 2 + (size_t)(g.S / 2) * 4; }

int generic_launch(const LaunchGeom& g, uint32_t metric, const uint32_t* bfly_labels_dev, const void* in,
                   uint8_t* out, int32_t* path_metric, uint32_t* surv, cudaStream_t st) {
    const size_t smem = generic_smem_bytes(g);
    if (metric == VDEC_HARD_U16) {
        auto k = vit_generic_kernel<HardOps, true>;
        if (smem > 48 * 1024) VDEC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<g.nframes, kGenThreads, smem, st>>>(g, bfly_labels_dev, in, out, path_metric, surv);
    } else {
        auto k = vit_generic_kernel<SoftOps, false>;
        if (smem > 48 * 1024) VDEC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<g.nframes, kGenThreads, smem, st>>>(g, bfly_labels_dev, in, out, path_metric, surv);
    }
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vd
