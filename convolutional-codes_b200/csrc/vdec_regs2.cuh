// vdec_regs2.cuh — integer-metric Viterbi kernel: ONE THREAD PER TWO FRAMES, packed 16-bit metrics.
//
// Serves (a) the hard-decision decoder (binary-symmetric-channel/viterbi-decoder.c: uint16 path
// metrics, Hamming branch metrics) and (b) the quantised soft mode VDEC_SOFT_U8 (SURVEY.md §8f-3:
// small non-negative integer branch metrics such as common/hard-demapper.c's {0,1,2}; on such inputs
// the reference's float decoder is exact integer arithmetic, so an integer kernel is bit-identical).
//
// Every 32-bit register holds the metric of the SAME state for TWO frames (low / high half), so one
// instruction updates two frames:
//     VIADD (32-bit add of packed halves; halves never carry: metrics < 2^15)
//     VIMNMX.U16x2  — packed min that also returns one predicate per half (sm_90+ DPX), i.e. min AND
//                     both decisions in a single issue slot
//     2 predicated adds drop the decision bits into the survivor word
// = 5 issue slots per TWO state updates, against 5 per ONE in the fp32 kernel.  Everything else
// (in-place butterflies, rotation period K-1, compile-time labels, tie rule, lowest-index argmin,
// survivor scratch + prefetched traceback, staged output) is as in vdec_regs.cuh.
//
// Input is consumed from a transposed staging buffer written by pack_kernel (vdec_regs2.cu):
//     tin[global warp][step][lane] = uint2 { frame 2*lane's step data, frame 2*lane+1's }
// so each warp-step is one coalesced 256-byte read.
//
// Range: unreachable = 0x8000, real metrics must stay below it: T * max_branch_metric < 0x8000 is
// checked on the host (the 32-bit kernels take over otherwise).  Below that bound the reference's
// clamp at 0xFF00 (viterbi-decoder.c:127-130) can never trigger, so decisions, final state and
// path_metric are identical; states still unreachable never win a compare against a real metric and
// are never visited by the traceback.
#pragma once
#include <type_traits>
#include <utility>

#include "vdec_regs.cuh"

namespace vd {

constexpr uint32_t kBig2 = 0x80008000u;

enum { kModeHard = 0, kModeQuant = 1 };

// branch metrics of one step for both frames, packed: bm[e] = mA(e) | mB(e) << 16
template <int M, int MODE>
__device__ __forceinline__ void packed_bm(uint2 x, uint32_t (&bm)[M]) {
    if constexpr (MODE == kModeHard) {
        const uint32_t r = (x.x & (M - 1)) | ((x.y & (M - 1)) << 16);        // :235-237, both frames
#pragma unroll
        for (int e = 0; e < M; ++e) {                                        // Hamming distance, n <= 3 bits (:69-72)
            const uint32_t d = r ^ (e * 0x00010001u);
            bm[e] = (d & 0x00010001u) + ((d >> 1) & 0x00010001u) + ((d >> 2) & 0x00010001u);
        }
    } else {
        static_assert(M == 4, "quantised mode: 4 metrics per step");
#pragma unroll
        for (int e = 0; e < M; ++e) bm[e] = __byte_perm(x.x, x.y, 0x4040 | e | ((4 + e) << 8)) & 0x00FF00FFu;   // (A.byte e, B.byte e)
    }
}

// packed min of two frames' candidates plus both decisions: min.u16x2 + the two setp.eq on the halves
// are what __vibmin_u16x2 expands to and what ptxas fuses into ONE VIMNMX.U16x2 with two predicate
// outputs; pred = (min == even) = (even <= odd), so the decision "odd strictly smaller" (tie -> even
// predecessor, viterbi-decoder.c:122) is its negation and ORs the frame's bit into the survivor word.
__device__ __forceinline__ uint32_t acs2(uint32_t ce, uint32_t co, uint32_t& dw, uint32_t bit_lo, uint32_t bit_hi) {
    uint32_t mn;
    asm("{\n\t.reg .pred pu, pv;\n\t.reg .u16 rs0, rs1, rs2, rs3;\n\t"
        "min.u16x2 %0, %2, %3;\n\t"
        "mov.b32 {rs0, rs1}, %0;\n\t"
        "mov.b32 {rs2, rs3}, %2;\n\t"
        "setp.eq.u16 pv, rs0, rs2;\n\t"
        "setp.eq.u16 pu, rs1, rs3;\n\t"
        "@!pv or.b32 %1, %1, %4;\n\t"
        "@!pu or.b32 %1, %1, %5;\n\t}"
        : "=r"(mn), "+r"(dw) : "r"(ce), "r"(co), "r"(bit_lo), "r"(bit_hi));
    return mn;
}

template <int S> struct Surv2 { static constexpr int NW = (S + 15) / 16; };   // 16 states x 2 frames per word

template <class CD, int PH>
__device__ __forceinline__ void acs_step2(uint32_t (&r)[CD::S], const uint32_t (&bm)[CD::M], uint32_t (&dw)[Surv2<CD::S>::NW]) {
    constexpr int B = CD::B, S = CD::S, PB = PH % B;
#pragma unroll
    for (int w = 0; w < Surv2<S>::NW; ++w) dw[w] = 0;
#pragma unroll
    for (int a = 0; a < S; ++a) {
        if ((a >> PB) & 1) continue;
        const int b = a | (1 << PB);
        const uint32_t u = rotr_bits(a, PH, B), v = u + 1;
        const uint32_t j0 = u >> 1, j1 = j0 + S / 2;
        const uint32_t pe = r[a], po = r[b];
        r[a] = acs2(pe + bm[CD::label(u, 0)], po + bm[CD::label(v, 0)], dw[j0 >> 4], 1u << (j0 & 15), 1u << (16 + (j0 & 15)));
        r[b] = acs2(pe + bm[CD::label(u, 1)], po + bm[CD::label(v, 1)], dw[j1 >> 4], 1u << (j1 & 15), 1u << (16 + (j1 & 15)));
    }
}

template <int NW> struct SurvVec;
template <> struct SurvVec<1> { using type = uint32_t; };
template <> struct SurvVec<2> { using type = uint2; };
template <> struct SurvVec<4> { using type = uint4; };

template <int NW>
__device__ __forceinline__ uint32_t surv_word(const typename SurvVec<NW>::type& v, uint32_t idx) {
    if constexpr (NW == 1) return v;
    else if constexpr (NW == 2) return idx ? v.y : v.x;
    else return (idx & 2) ? ((idx & 1) ? v.w : v.z) : ((idx & 1) ? v.y : v.x);
}

constexpr int kTrace2Chunk = 8;

template <class CD, int MODE>
__global__ void __launch_bounds__(kRegsThreads, 4)
vit_regs2_kernel(const uint2* __restrict__ tin, uint8_t* __restrict__ out, int32_t* __restrict__ path_metric,
                 uint32_t* __restrict__ surv, uint32_t nframes, uint32_t T, uint32_t out_bytes, uint32_t stage_words,
                 uint32_t init_all_from) {
    constexpr int S = CD::S, B = CD::B, M = CD::M, NW = Surv2<S>::NW;
    using SV = typename SurvVec<NW>::type;
    extern __shared__ __align__(16) unsigned char s_raw[];      // output staging: [warp][64 rows][stage_words]

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gt = blockIdx.x * kRegsThreads + tid;        // thread <-> frames 2*gt, 2*gt+1
    const uint32_t gw = gt >> 5;
    const uint32_t fA = 2 * gt, fB = fA + 1;
    const bool liveA = fA < nframes, liveB = fB < nframes;
    const uint2* src = tin + (size_t)gw * T * 32 + lane;
    SV* sv = reinterpret_cast<SV*>(surv) + (size_t)gw * T * 32 + lane;

    uint32_t r[S];
    r[0] = 0;                                                   // :225
    const uint32_t rest = (fA >= init_all_from ? 0u : 0x8000u) | (fB >= init_all_from ? 0u : 0x80000000u);
#pragma unroll
    for (int s = 1; s < S; ++s) r[s] = rest;                    // unreachable (:226-228) unless the start is unknown

    uint32_t dw[NW], bm[M];
    uint2 xn = __ldg(src);
    auto store_dw = [&](uint32_t t) {
        SV v;
        if constexpr (NW == 1) v = dw[0];
        else if constexpr (NW == 2) v = make_uint2(dw[0], dw[1]);
        else v = make_uint4(dw[0], dw[1], dw[2], dw[3]);
        if (liveA) sv[(size_t)t * 32] = v;
    };

    uint32_t t = 0;
    for (; t + B <= T; t += B) {
        static_for<0, B>([&](auto phc) {
            constexpr int PH = decltype(phc)::value;
            const uint2 x = xn;
            xn = __ldg(src + (size_t)min(t + PH + 1, T - 1) * 32);           // one step of lookahead, coalesced
            packed_bm<M, MODE>(x, bm);
            acs_step2<CD, PH>(r, bm, dw);
            store_dw(t + PH);
        });
    }
    const uint32_t rem = T - t;
    static_for<0, B - 1>([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (PH < (int)rem) {
            const uint2 x = xn;
            xn = __ldg(src + (size_t)min(t + PH + 1, T - 1) * 32);
            packed_bm<M, MODE>(x, bm);
            acs_step2<CD, PH>(r, bm, dw);
            store_dw(t + PH);
        }
    });

    // final states: lowest-index argmin per frame, strict '<' from the sentinel (:75-82)
    uint32_t bestA = 0xFF00u, bestB = 0xFF00u, curA = 0, curB = 0;
    static_for<0, B>([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (PH == (int)rem) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const uint32_t v = r[rotl_bits(s, PH, B)];
                const uint32_t a = v & 0xFFFFu, b = v >> 16;
                if (a < bestA) { bestA = a; curA = s; }
                if (b < bestB) { bestB = b; curB = s; }
            }
        }
    });
    if (path_metric != nullptr) {                                                             // :94, :143
        if (liveA) path_metric[fA] = (int32_t)bestA;
        if (liveB) path_metric[fB] = (int32_t)bestB;
    }

    // traceback of both frames (:84-93); survivor vectors fetched a chunk ahead
    uint32_t* stage = reinterpret_cast<uint32_t*>(s_raw) + ((size_t)warp * 64 + 2 * lane) * stage_words;
    if (liveA) {
        SV wc[kTrace2Chunk], wn[kTrace2Chunk];
        int base = ((int)T - 1) & ~(kTrace2Chunk - 1);
#pragma unroll
        for (int j = 0; j < kTrace2Chunk; ++j)
            if (base + j < (int)T) wc[j] = __ldcg(sv + (size_t)(base + j) * 32);
        uint32_t accA = 0, accB = 0;
        for (; base >= 0; base -= kTrace2Chunk) {
            if (base >= kTrace2Chunk) {
#pragma unroll
                for (int j = 0; j < kTrace2Chunk; ++j) wn[j] = __ldcg(sv + (size_t)(base - kTrace2Chunk + j) * 32);
            }
#pragma unroll
            for (int j = kTrace2Chunk - 1; j >= 0; --j) {
                const int tt = base + j;
                if (tt < (int)T) {
                    const uint32_t dA = (surv_word<NW>(wc[j], curA >> 4) >> (curA & 15)) & 1u;
                    const uint32_t dB = (surv_word<NW>(wc[j], curB >> 4) >> (16 + (curB & 15))) & 1u;
                    accA |= (curA >> (B - 1)) << (31 - (tt & 31));
                    accB |= (curB >> (B - 1)) << (31 - (tt & 31));
                    curA = ((curA << 1) & (S - 1)) | dA;
                    curB = ((curB << 1) & (S - 1)) | dB;
                    if ((tt & 31) == 0) {
                        stage[tt >> 5] = __byte_perm(accA, 0, 0x0123);
                        stage[stage_words + (tt >> 5)] = __byte_perm(accB, 0, 0x0123);
                        accA = 0; accB = 0;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kTrace2Chunk; ++j) wc[j] = wn[j];
        }
    }
    __syncwarp();

    // the warp's 64 output rows are contiguous in global memory
    const uint32_t wbase = 2 * (blockIdx.x * kRegsThreads + (tid & ~31u));
    if (wbase < nframes) {
        const uint32_t rows = min(64u, nframes - wbase);
        const uint32_t total = rows * out_bytes;
        const uint8_t* sb = s_raw + (size_t)warp * 64 * stage_words * 4;
        uint8_t* dst = out + (size_t)wbase * out_bytes;
        uint32_t row = 0, col = lane;
        while (col >= out_bytes) { col -= out_bytes; ++row; }
        for (uint32_t i = lane; i < total; i += 32) {
            dst[i] = sb[(size_t)row * stage_words * 4 + col];
            col += 32;
            while (col >= out_bytes) { col -= out_bytes; ++row; }
        }
    }
}

// ---- transposition of the caller's rows into the kernel's coalesced staging layout ---------------
// hard: in = uint8 [F][T] (one symbol per step); quant: in = uint8 [F][T][4]
// One CTA = one warp-group of 64 frames x kPackSteps steps: rows are read along t (coalesced), parked in
// shared memory, and written out so that the 32 lanes' uint2 of one step form one 256-byte line.
constexpr int kPackSteps = 64;
template <int MODE>
__global__ void __launch_bounds__(256)
pack2_kernel(const uint8_t* __restrict__ in, uint2* __restrict__ tin, uint32_t nframes, uint32_t T, uint32_t stride) {
    __shared__ uint32_t tile[64][kPackSteps + 1];
    const uint32_t gw = blockIdx.y, t0 = blockIdx.x * kPackSteps, tid = threadIdx.x;
    const uint32_t steps = min((uint32_t)kPackSteps, T - t0);
    for (uint32_t i = tid; i < 64 * kPackSteps; i += 256) {
        const uint32_t row = i / kPackSteps, c = i % kPackSteps;
        uint32_t f = gw * 64 + row;
        if (f >= nframes) f = nframes - 1;                       // an odd tail shadows the last frame
        if (c < steps) {
            if constexpr (MODE == kModeHard) tile[row][c] = in[(size_t)f * stride + t0 + c];
            else tile[row][c] = __ldg(reinterpret_cast<const uint32_t*>(in) + (size_t)f * stride + t0 + c);
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < 32 * steps; i += 256) {
        const uint32_t c = i >> 5, lane = i & 31;
        tin[((size_t)gw * T + t0 + c) * 32 + lane] = make_uint2(tile[2 * lane][c], tile[2 * lane + 1][c]);
    }
}

struct Regs2Entry {
    const char* name;
    uint32_t K, n, mode;
    uint8_t labels[128];
    int (*launch)(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch, cudaStream_t st);
    size_t (*scratch_bytes)(const LaunchGeom& g);
};
const Regs2Entry* regs2_find(uint32_t K, uint32_t n, uint32_t mode, const uint8_t* labels);

template <class CD>
size_t regs2_scratch_bytes(const LaunchGeom& g) {
    const size_t warps = ((size_t)(g.nframes + 1) / 2 + 31) / 32;
    return warps * g.T * 32 * (sizeof(uint2) + Surv2<CD::S>::NW * sizeof(uint32_t));    // staging + survivors
}

template <class CD, int MODE>
int regs2_launch(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch, cudaStream_t st) {
    const size_t npairs = ((size_t)g.nframes + 1) / 2, warps = (npairs + 31) / 32;
    uint2* tin = reinterpret_cast<uint2*>(scratch);
    uint32_t* surv = reinterpret_cast<uint32_t*>(tin + warps * g.T * 32);
    pack2_kernel<MODE><<<dim3((g.T + kPackSteps - 1) / kPackSteps, (unsigned)warps), 256, 0, st>>>(
        static_cast<const uint8_t*>(in), tin, g.nframes, g.T, g.stride_steps);
    VDEC_CUDA(cudaGetLastError());
    uint32_t stage_words = (g.out_bytes + 3) / 4;
    if ((stage_words & 1) == 0) ++stage_words;
    const size_t smem = (size_t)(kRegsThreads / 32) * 64 * stage_words * 4;
    auto kern = vit_regs2_kernel<CD, MODE>;
    if (smem > 48 * 1024) VDEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (uint32_t)((npairs + kRegsThreads - 1) / kRegsThreads);
    kern<<<blocks, kRegsThreads, smem, st>>>(tin, out, path_metric, surv, g.nframes, g.T, g.out_bytes, stage_words, g.init_all_from);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

template <class CD, int MODE>
Regs2Entry make_regs2_entry(const char* name) {
    Regs2Entry e{};
    e.name = name; e.K = CD::K; e.n = CD::N; e.mode = MODE;
    for (int s = 0; s < CD::S; ++s) { e.labels[2 * s] = (uint8_t)CD::label(s, 0); e.labels[2 * s + 1] = (uint8_t)CD::label(s, 1); }
    e.launch = &regs2_launch<CD, MODE>;
    e.scratch_bytes = &regs2_scratch_bytes<CD>;
    return e;
}

}  // namespace vd
