// vdec_tile2.cuh — register-tile kernel with PACKED 16-bit integer metrics for K = 8 .. 15: the large-trellis
// counterpart of vdec_regs2.cuh.
//
// Serves the hard-decision decoder (binary-symmetric-channel/viterbi-decoder.c:38-150: uint16 path metrics, Hamming
// branch metrics) and the quantised soft mode (VDEC_SOFT_U8) on trellises that do not fit one thread.  Structure =
// vdec_tile.cuh (2^(K-7) threads per frame PAIR, 64 metrics per thread, windows of 6 local steps, shared-memory
// transposes sigma' = rotr(sigma, 6) through the same conflict-free map phi), arithmetic = vdec_regs2.cuh (every
// register holds the metric of one state for TWO frames; per butterfly 2 adds, 2 VIADDMNMX.U16x2, 2 three-input
// IADD3 whose bits 15 / 31 are the strict-'<' decisions, 1 PRMT and 1 LOP3 that gather them: 2 issue slots per state
// update).  Range condition as there: T * max branch metric < 0x4000, checked on the host.
//
// Labels.  True parity: label = (register part, compile time) ^ (thread part xt); hard symbols absorb it for free
// (popc(e ^ xt ^ r) = Hamming distance of e to r ^ xt), quantised metrics by one byte permute of the step's word.
// De-facto labels of the reference (vdec_labels.h): the 4-bit linear extended label of vdec_tile.cuh; the window's
// packed metrics are expanded to 16 entries per step in shared memory by the pair's threads and fetched by label.
//
// Decisions: 4 words per thread per step (64 registers x 2 frames), [pair][step][thread] uint4 in HBM, bit layout of
// vdec_regs2.cuh inside a thread (butterfly m of the thread, input h, frame fr: word m / 8, bit 8 (2h + fr) + m % 8).
// Traceback: one thread per frame, one window (<= 6 steps, same owner thread) of survivor vectors fetched at a time.
#pragma once
#include "vdec_regs2.cuh"
#include "vdec_tile.cuh"

namespace vd {

// does any of the 64 registers carry (extended) label E at this step position?
template <class CD>
VDEC_HD constexpr bool label_used(uint32_t E) {
    for (uint32_t u = 0; u < 64; ++u)
        for (uint32_t i = 0; i < 2; ++i) if (CD::label(u, i) == E) return true;
    return false;
}

template <class TD, int MODE>
__global__ void __launch_bounds__(kTileThreads, 2)
vit_tile2_forward(const void* __restrict__ in, uint4* __restrict__ surv, uint32_t* __restrict__ final_state,
                  int32_t* __restrict__ path_metric, uint32_t nframes, uint32_t T, uint32_t stride_steps,
                  uint32_t init_all_from) {
    constexpr int B = TD::B, S = TD::S, TB = TD::TB, TPF = TD::TPF, PPC = TD::FPC, RT = TD::RT, M = TD::M;
    constexpr int PPW = TPF < 32 ? 32 / TPF : 1, PSH = ilog2(PPW);          // frame pairs per warp
    extern __shared__ __align__(1024) unsigned char s_raw[];
    uint32_t* xall = reinterpret_cast<uint32_t*>(s_raw);                        // [NREG][RT * 64] exchange regions
    unsigned char* vms = s_raw + (size_t)kTileThreads * 64 * 4;                // REF: [warp or pair][6][16][PPW] packed metrics
    uint32_t* red = reinterpret_cast<uint32_t*>(vms + (TD::REF ? (size_t)PPC * kTileVmBytes : 0));   // [2 frames][kTileThreads / 32][2]

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t tau = tid & (TPF - 1), pic = tid / TPF;                       // thread of the pair, pair in the CTA
    const uint32_t pair = blockIdx.x * PPC + pic;
    const uint32_t fA = 2 * pair, fB = fA + 1;
    const bool liveA = fA < nframes, liveB = fB < nframes;
    const uint32_t thr = tid & (RT - 1);
    uint32_t* xbuf = xall + (size_t)(tid / RT) * RT * 64;
    const uint32_t wbase = TD::wthr(thr), rbase = TD::rthr(thr);

    uint32_t xt[6];                                      // thread part of the (extended) labels per step position
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < TB; ++j) x ^= ((tau >> j) & 1u) ? TD::ebit(6 + j - k) : 0u;
        xt[k] = x;
    }

    // rows of the two frames; a dead frame shadows the last live one and stores nothing
    using Elem = std::conditional_t<MODE == kModeHard, uint8_t, uint32_t>;      // one step: 1 symbol byte / 4 metric bytes
    const Elem* rowA = static_cast<const Elem*>(in) + (size_t)min(fA, nframes - 1) * stride_steps;
    const Elem* rowB = static_cast<const Elem*>(in) + (size_t)min(fB, nframes - 1) * stride_steps;
    auto fetch = [&](uint32_t t) { return make_uint2((uint32_t)__ldg(rowA + t), (uint32_t)__ldg(rowB + t)); };

    [[maybe_unused]] const uint32_t piw = pic & (PPW - 1);
    [[maybe_unused]] uint32_t* vm = reinterpret_cast<uint32_t*>(vms + (size_t)(pic >> PSH) * PPW * kTileVmBytes) + piw;
    [[maybe_unused]] const uint32_t vma = (uint32_t)__cvta_generic_to_shared(vm);

    // decoder_reset (:222-232): state 0 starts at 0, the rest unreachable — unless the start state is unknown
    const uint32_t rest = (fA >= init_all_from ? 0u : (kBig2 & 0xFFFFu)) | (fB >= init_all_from ? 0u : (kBig2 & 0xFFFF0000u));
    uint32_t r[64];
#pragma unroll
    for (int a = 0; a < 64; ++a) r[a] = rest;
    if (tau == 0) r[0] = 0;

    uint4* sv = surv + (size_t)pair * T * TPF + tau;
    uint32_t t = 0;
    [[maybe_unused]] uint2 xn = fetch(0);
    auto begin_window = [&]() {
        tile_group_sync<TPF>(pic);                      // the previous exchange (and expansion) has been consumed by all
        if constexpr (TD::REF) {                        // vm[step][el] = packed metric of label (el & 3) & ~(el >> 2)
#pragma unroll
            for (int e = 0; e < (96 + TPF - 1) / TPF; ++e) {
                const uint32_t i = e * TPF + tau;
                if (TPF >= 96 ? tau < 96u : i < 96u) {
                    uint32_t bm[M];
                    packed_bm<M, MODE>(fetch(min(t + (i >> 4), T - 1)), bm);
                    const uint32_t mi = TD::metric_index(i & 15u);
                    vm[(size_t)i << PSH] = mi == 0 ? bm[0] : mi == 1 ? bm[1] : mi == 2 ? bm[2] : bm[3];
                }
            }
            tile_group_sync<TPF>(pic);
        }
    };
    auto step = [&](auto kc) {
        constexpr int KS = decltype(kc)::value;
        using CD = TileCode<TD, KS>;
        uint32_t bm[CD::M], dw[4];
        if constexpr (TD::REF) {
            const uint32_t bx = (vma + KS * (64 << PSH)) ^ (xt[KS] << (2 + PSH));
            static_for<0, 16>([&](auto ec) {
                constexpr uint32_t E = decltype(ec)::value;
                if constexpr (label_used<CD>(E)) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(bm[E]) : "r"(bx ^ (E << (2 + PSH))));
                else bm[E] = 0;
            });
        } else {
            uint2 x = xn;
            xn = fetch(min(t + KS + 1, T - 1));                                     // one step of lookahead
            if constexpr (MODE == kModeHard) { x.x ^= xt[KS]; x.y ^= xt[KS]; }       // Hamming distance to e ^ xt = distance of e to r ^ xt
            else {                                                                   // metric of label e ^ xt: permute the 4 bytes
                const uint32_t sel = 0x3210u ^ (xt[KS] * 0x1111u);
                x.x = __byte_perm(x.x, 0, sel); x.y = __byte_perm(x.y, 0, sel);
            }
            packed_bm<M, MODE>(x, bm);
        }
        acs_step2<CD, KS>(r, bm, dw);
        if (liveA) sv[(size_t)(t + KS) * TPF] = make_uint4(dw[0], dw[1], dw[2], dw[3]);
    };
    for (; t + 6 < T; t += 6) {
        begin_window();
        static_for<0, 6>(step);
        // ---- transpose: sigma' = rotr(sigma, 6) (address map and conflict-freeness: TileDef::phi) -----------------
        static_for<0, 64>([&](auto ac) {
            constexpr uint32_t A = decltype(ac)::value, c = TD::cw(A), cx = c & TD::wmask(), cd = c & ~TD::wmask();
            xbuf[(wbase ^ cx) + cd] = r[A];
        });
        tile_group_sync<TPF>(pic);
        static_for<0, 64 / TD::VEC>([&](auto vc) {
            constexpr uint32_t V = decltype(vc)::value, c = TD::cr(TD::VEC * V), cx = c & TD::rmask(), cd = c & ~TD::rmask();
            const uint32_t* src = xbuf + ((rbase ^ cx) + cd);
            if constexpr (TD::VEC == 4) {
                const uint4 q = *reinterpret_cast<const uint4*>(src);
                r[4 * V] = q.x; r[4 * V + 1] = q.y; r[4 * V + 2] = q.z; r[4 * V + 3] = q.w;
            } else if constexpr (TD::VEC == 2) {
                const uint2 q = *reinterpret_cast<const uint2*>(src);
                r[2 * V] = q.x; r[2 * V + 1] = q.y;
            } else {
                r[V] = *src;
            }
        });
    }
    const uint32_t n = T - t;                          // last window: 1 .. 6 steps
    begin_window();
    static_for<0, 6>([&](auto kc) { if ((int)n > (decltype(kc)::value)) step(kc); });

    // ---- final states: lowest LOGICAL index among the minima, strict '<' from the 0xFF00 sentinel (:75-82) --------
    uint32_t bestA = 0xFF00u, bestB = 0xFF00u, sA = 0xFFFFFFFFu, sB = 0xFFFFFFFFu;
    static_for<0, 6>([&](auto kc) {
        constexpr int KS = decltype(kc)::value;
        if (KS + 1 == (int)n) {
            const uint32_t x = tau << 6;
            const uint32_t st = ((x >> (KS + 1)) | (x << (B - KS - 1))) & (S - 1);
#pragma unroll
            for (int a = 0; a < 64; ++a) {
                const uint32_t va = r[a] & 0xFFFFu, vb = r[a] >> 16, s = st | rotr_bits((uint32_t)a, KS + 1, B);
                if (va < bestA || (va == bestA && s < sA)) { bestA = va; sA = s; }
                if (vb < bestB || (vb == bestB && s < sB)) { bestB = vb; sB = s; }
            }
        }
    });
    auto better = [](uint32_t v, uint32_t s, uint32_t bv, uint32_t bs) { return v < bv || (v == bv && s < bs); };
#pragma unroll
    for (int off = 1; off < (TPF < 32 ? TPF : 32); off <<= 1) {
        const uint32_t oa = __shfl_xor_sync(0xffffffffu, bestA, off), osa = __shfl_xor_sync(0xffffffffu, sA, off);
        const uint32_t ob = __shfl_xor_sync(0xffffffffu, bestB, off), osb = __shfl_xor_sync(0xffffffffu, sB, off);
        if (better(oa, osa, bestA, sA)) { bestA = oa; sA = osa; }
        if (better(ob, osb, bestB, sB)) { bestB = ob; sB = osb; }
    }
    if constexpr (TPF > 32) {
        uint32_t* slot = red + (tid >> 5) * 4;
        if (lane == 0) { slot[0] = bestA; slot[1] = sA; slot[2] = bestB; slot[3] = sB; }
        __syncthreads();
        if (tau == 0) {
#pragma unroll
            for (int w = 1; w < TPF / 32; ++w) {
                const uint32_t* o = slot + w * 4;
                if (better(o[0], o[1], bestA, sA)) { bestA = o[0]; sA = o[1]; }
                if (better(o[2], o[3], bestB, sB)) { bestB = o[2]; sB = o[3]; }
            }
        }
    }
    if (tau == 0) {
        if (liveA) { final_state[fA] = sA == 0xFFFFFFFFu ? 0u : sA; if (path_metric) path_metric[fA] = (int32_t)bestA; }   // :94, :143
        if (liveB) { final_state[fB] = sB == 0xFFFFFFFFu ? 0u : sB; if (path_metric) path_metric[fB] = (int32_t)bestB; }
    }
}

static __global__ void __launch_bounds__(128)
vit_tile2_traceback(const uint4* __restrict__ surv, const uint32_t* __restrict__ final_state, uint8_t* __restrict__ out,
                    uint32_t nframes, uint32_t T, uint32_t out_bytes, uint32_t B, uint32_t TPF) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const uint4* sv = surv + (size_t)(f >> 1) * T * TPF;
    const uint32_t fr = f & 1u;
    uint8_t* row = out + (size_t)f * out_bytes;
    const uint32_t S1 = (1u << B) - 1u;
    uint32_t h = final_state[f];
    for (int w0 = (int)((T - 1) / 6) * 6; w0 >= 0; w0 -= 6) {
        const int n = (int)T - w0 < 6 ? (int)T - w0 : 6;
        const uint32_t owner = ((h & S1) >> (6 - n)) & (TPF - 1);
        uint4 wd[6];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (k < n) wd[k] = __ldcg(sv + (size_t)(w0 + k) * TPF + owner);
#pragma unroll
        for (int k = 5; k >= 0; --k) {
            if (k >= n) continue;
            const int tt = w0 + k;
            if ((tt & 15) == 0) {
                const uint32_t rv = __brev(h >> (B - 1));
                row[tt >> 3] = (uint8_t)(rv >> 24);
                if (tt + 8 < (int)T) row[(tt >> 3) + 1] = (uint8_t)(rv >> 16);
            }
            const uint32_t cur = h & S1;
            const uint32_t j = (cur & ((1u << (5 - k)) - 1u)) | ((cur >> (B - 1 - k)) << (5 - k));   // local index of the state in its thread
            const uint32_t m = j & 31u, wi = m >> 3;
            const uint32_t word = (wi & 2u) ? ((wi & 1u) ? wd[k].w : wd[k].z) : ((wi & 1u) ? wd[k].y : wd[k].x);
            h = (h << 1) | ((word >> (8 * (2 * (j >> 5) + fr) + (m & 7u))) & 1u);                    // predecessor (:90)
        }
    }
}

#ifndef __CUDACC_RTC__
inline size_t tile2_surv_bytes_tpf(const LaunchGeom& g, uint32_t tpf) {
    const size_t pairs = ((size_t)g.nframes + 1) / 2;
    return pairs * g.T * tpf * sizeof(uint4) + (size_t)g.nframes * sizeof(uint32_t);
}

// `kern`: vit_tile2_forward<TD, MODE> of this library or a cudaKernel_t of a run-time compiled module (vdec_jit.cu)
inline int tile2_launch_kernel(const void* kern, const TileShape& sh, const LaunchGeom& g, const void* in, uint8_t* out,
                               int32_t* path_metric, uint32_t* scratch, cudaStream_t st) {
    const size_t pairs = ((size_t)g.nframes + 1) / 2;
    uint4* surv = reinterpret_cast<uint4*>(scratch);
    uint32_t* final_state = reinterpret_cast<uint32_t*>(surv + pairs * g.T * sh.TPF);
    const size_t smem = (size_t)kTileThreads * 64 * 4 + (sh.ref ? (size_t)sh.FPC * kTileVmBytes : 0) + (kTileThreads / 32) * 16;
    VDEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (uint32_t)((pairs + sh.FPC - 1) / sh.FPC);
    uint32_t nframes = g.nframes, T = g.T, stride = g.stride_steps, init_all_from = g.init_all_from;
    void* args[] = {&in, &surv, &final_state, &path_metric, &nframes, &T, &stride, &init_all_from};
    VDEC_CUDA(cudaLaunchKernel(kern, dim3(blocks), dim3(kTileThreads), args, smem, st));
    vit_tile2_traceback<<<(g.nframes + 127) / 128, 128, 0, st>>>(surv, final_state, out, g.nframes, g.T, g.out_bytes, sh.B, sh.TPF);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

template <class TD>
size_t tile2_surv_bytes(const LaunchGeom& g) { return tile2_surv_bytes_tpf(g, TD::TPF); }

template <class TD, int MODE>
int tile2_launch(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch, cudaStream_t st) {
    return tile2_launch_kernel(reinterpret_cast<const void*>(&vit_tile2_forward<TD, MODE>),
                               TileShape{(uint32_t)TD::B, (uint32_t)TD::TPF, (uint32_t)TD::FPC, TD::REF, false}, g, in, out, path_metric,
                               scratch, st);
}

struct Tile2Entry {
    const char* name;
    uint32_t K, n, mode;
    const uint8_t* labels;                     // [S][2]
    int (*launch)(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch, cudaStream_t st);
    size_t (*bytes)(const LaunchGeom& g);
};

template <class TD, int MODE>
Tile2Entry make_tile2_entry(const char* name) {
    return Tile2Entry{name, (uint32_t)TD::K, (uint32_t)TD::N, (uint32_t)MODE, tile_labels<TD>(), &tile2_launch<TD, MODE>, &tile2_surv_bytes<TD>};
}
#endif  // !__CUDACC_RTC__

}  // namespace vd
