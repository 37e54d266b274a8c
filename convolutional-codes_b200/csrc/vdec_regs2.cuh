// vdec_regs2.cuh — integer-metric Viterbi kernel: ONE THREAD PER TWO FRAMES, packed 16-bit metrics.
//
// Serves (a) the hard-decision decoder (binary-symmetric-channel/viterbi-decoder.c: uint16 path
// metrics, Hamming branch metrics) and (b) the quantised soft mode VDEC_SOFT_U8 (SURVEY.md §8f-3:
// small non-negative integer branch metrics such as common/hard-demapper.c's {0,1,2}; on such inputs
// the reference's float decoder is exact integer arithmetic, so an integer kernel is bit-identical).
//
// Every 32-bit register holds the metric of the SAME state for TWO frames (low / high half), so one
// instruction updates two frames.  Per butterfly (2 states x 2 frames = 4 state updates), 8 issue slots:
//     2 x add            ce = pe + bm[even label]              (halves never carry: every metric < 2^15)
//     2 x VIADDMNMX.U16x2  new = min(po + bm[odd label], ce)   (sm_90+ DPX: add and packed min in one slot)
//     2 x IADD3          t  = ce - po + (0x7FFF7FFF - bm[odd]) = ce - co + 0x7FFF per half (for S >= 16 formed as
//                        (pe - po) + table[label pair] with three FMA-pipe adds, see acs_step2): bit 15 of each
//                        half is set iff co < ce, i.e. the reference's strict '<' decision (tie -> even
//                        predecessor, viterbi-decoder.c:122), with no borrow between the halves
//     1 x PRMT           sign-replicates the four decision bytes of (t0, t1) into 0x00 / 0xFF bytes
//     1 x LOP3           survivor word |= bytes & (0x01010101 << slot)
// = 2 issue slots per state update and no predicate traffic, against ~4.4 in the fp32 kernel.  Measured
// (tools/ubench/pipes.cu, profiles/r01l): VIADDMNMX, 3-input IADD3, PRMT and LOP3 all issue to the half-rate
// ALU pipe, so the bound is 6 ALU slots = 12 clk per butterfly = 3 clk per state update (193 Gbit/s at K=7).  Everything
// else (in-place butterflies, rotation period K-1, compile-time labels, tie rule, lowest-index argmin,
// survivor scratch + prefetched traceback, staged output) is as in vdec_regs.cuh.
//
// Input: every lane reads the rows of its two frames directly, one step (1 byte hard / 4 bytes quantised)
// ahead of the ACS; a 32-byte sector serves 32 / 8 consecutive steps of a row and stays in L1 meanwhile
// (128 threads x 2 rows x 4 CTAs x 32 B = 32 KB per SM), so no transposing pre-pass is needed.
//
// Range: unreachable = 0x4000 and every metric must stay below 0x8000 (so that differences of two
// metrics fit 16 bits signed): T * max_branch_metric < 0x4000 is checked on the host (the 32-bit kernels
// take over otherwise).  Below that bound the reference's clamp at 0xFF00 (viterbi-decoder.c:127-130) can
// never trigger on a reachable state, so decisions along every survivor the traceback can visit, the
// final state and path_metric are identical; states still unreachable never win a compare against a real
// metric and are never visited by the traceback.
//
// Survivor word layout: butterfly m = j mod S/2 of new state j, h = top bit of j, frame fr in {0,1}:
// word m / 8, bit 8 * (2h + fr) + m % 8.
#pragma once
#include <type_traits>
#include <utility>

#include "vdec_regs.cuh"

namespace vd {

constexpr uint32_t kBig2 = 0x40004000u;       // unreachable, both halves
constexpr uint32_t kBias2 = 0x7FFF7FFFu;
constexpr uint32_t kRegs2Limit = 0x4000u;     // T * max branch metric must stay below this

enum { kModeHard = 0, kModeQuant = 1 };

// Hard decisions: the M packed Hamming metrics of a step depend on the two received n-bit symbols only, so they come
// out of a shared-memory table indexed by (symbol A, symbol B) — one LDS.128 (two for n = 3) instead of 16 - 28 ALU-pipe
// operations per step.  That is 6 % of a K = 7 step but 30 - 40 % of a step on the 4- and 8-state trellises of the
// codebook, where the per-step overhead is shared by very few butterflies (profiles/r02q_ab_hard_lut.txt).
#ifndef VDEC_HARD_LUT
#define VDEC_HARD_LUT 1
#endif
template <int N> constexpr int kHardLutWords = (1 << (2 * N)) * (1 << N);
template <int N>
__device__ __forceinline__ void hard_lut_init(uint32_t* lut) {
    constexpr int M = 1 << N;
    for (uint32_t i = threadIdx.x; i < (uint32_t)kHardLutWords<N>; i += blockDim.x) {
        const uint32_t e = i & (M - 1), ra = (i >> N) & (M - 1), rb = i >> (2 * N);
        lut[i] = (uint32_t)__popc(e ^ ra) | ((uint32_t)__popc(e ^ rb) << 16);          // :69-72 for both frames
    }
    __syncthreads();
}
template <int N>
__device__ __forceinline__ void hard_lut_bm(const uint32_t* lut, uint2 x, uint32_t (&bm)[1 << N]) {
    constexpr int M = 1 << N;
    const uint32_t idx = (x.x & (M - 1)) | ((x.y & (M - 1)) << N);                       // :235-237, both frames
    const uint4* row = reinterpret_cast<const uint4*>(lut + idx * M);
    if constexpr (M == 2) { const uint2 v = *reinterpret_cast<const uint2*>(lut + idx * 2); bm[0] = v.x; bm[1] = v.y; }
    else {
#pragma unroll
        for (int q = 0; q < M / 4; ++q) { const uint4 v = row[q]; bm[4 * q] = v.x; bm[4 * q + 1] = v.y; bm[4 * q + 2] = v.z; bm[4 * q + 3] = v.w; }
    }
}

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

#ifndef VDEC_REGS2_SPLIT
#define VDEC_REGS2_SPLIT 1
#endif
#ifndef VDEC_R2_PF_BYTES
#define VDEC_R2_PF_BYTES 32       // input rows: L1 prefetch distance in bytes (0 = none), see prefetch_rows
#endif
#ifndef VDEC_R2_MINB_LIN
#define VDEC_R2_MINB_LIN 3        // CTAs per SM of the 64-state kernels with true-parity labels: 16 distinct label pairs spill at 128 registers
#endif
// 64-state kernels whose labelling uses all M * M label pairs (true-parity labels, and rate 1/3 with either labelling):
// the table of pair constants does not fit next to 64 metric registers at 128 registers per thread
template <class CD> constexpr bool kRegs2Wide = CD::S >= 64 && (!CD::REF || CD::N == 3);
template <class CD> constexpr int kRegs2MinBlocks = kRegs2Wide<CD> ? VDEC_R2_MINB_LIN : 4;
// a - b that ptxas does not merge with a following add into one 3-input IADD3
__device__ __forceinline__ uint32_t sub_opaque(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ uint32_t add_opaque(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// a + b written as a * 1 + b: ptxas keeps it an IMAD (FMA pipes) instead of an ALU-pipe IADD3
__device__ __forceinline__ uint32_t mad1(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// does any butterfly of the code compare label pair I = (even label) * M + (odd label)?  (the same set at every phase)
template <class CD>
VDEC_HD constexpr bool label_pair_used(int I) {
    for (uint32_t u = 0; u < (uint32_t)CD::S; u += 2)
        for (uint32_t i = 0; i < 2; ++i)
            if ((int)(CD::label(u, i) * CD::M + CD::label(u + 1, i)) == I) return true;
    return false;
}

template <int S> struct Surv2 { static constexpr int NW = (S + 15) / 16; };   // 8 butterflies x 2 states x 2 frames per word

template <class CD, int PH>
__device__ __forceinline__ void acs_step2(uint32_t (&r)[CD::S], const uint32_t (&bm)[CD::M], uint32_t (&dw)[Surv2<CD::S>::NW]) {
    constexpr int B = CD::B, S = CD::S, PB = PH % B, M = CD::M;
    // Two forms of the decision word t (bit 15 of each half = strict '<'), equal modulo 2^32:
    //   fused  t = ce - po + kb[odd label]                    one 3-input IADD3 per successor: ALU pipe only
    //   split  t = (pe - po) + cc[even label][odd label]      the difference is shared by both successors and the
    //          constants cc = bm[even] - bm[odd] + 0x7FFF (per half) by all butterflies with that label pair: three
    //          2-input adds per butterfly, written as IMAD so that they go to the FMA pipes.
    // The kernel is bound by the half-rate ALU pipe (6 of its 8 slots per butterfly); the split form leaves 4 there
    // (measured K = 7 hard 131.7 -> 137.2 Gbit/s, quantised 128.0 -> 132.0, K = 15 hard 0.530 -> 0.553).  With few
    // butterflies per step the per-step table costs more than it saves (K = 3: 629 -> 594), hence S >= 16.
    constexpr bool kSplit = VDEC_REGS2_SPLIT && S >= 16;
    [[maybe_unused]] uint32_t cc[M * M];
    [[maybe_unused]] uint32_t kb[M];
    if constexpr (kSplit) {
        static_for<0, M * M>([&](auto ic) {
            constexpr int I = decltype(ic)::value;
            if constexpr (label_pair_used<CD>(I)) cc[I] = add_opaque(bm[I / M], kBias2 - bm[I % M]);
            else cc[I] = 0;
        });
    } else {
#pragma unroll
        for (int e = 0; e < M; ++e) kb[e] = kBias2 - bm[e];
    }
#pragma unroll
    for (int w = 0; w < Surv2<S>::NW; ++w) dw[w] = 0;
    static_for<0, S>([&](auto ac) {
        constexpr int a = decltype(ac)::value;
        if constexpr (((a >> PB) & 1) == 0) {
            constexpr int b = a | (1 << PB);
            constexpr uint32_t u = rotr_bits(a, PH, B), v = u + 1, m = u >> 1;     // new states m (input 0) and m + S/2 (input 1)
            const uint32_t pe = r[a], po = r[b];
            const uint32_t ce0 = pe + bm[CD::label(u, 0)], ce1 = pe + bm[CD::label(u, 1)];
            uint32_t t0, t1;
            if constexpr (kSplit) {
                const uint32_t dd = sub_opaque(pe, po);
                t0 = mad1(dd, cc[CD::label(u, 0) * M + CD::label(v, 0)]);
                t1 = mad1(dd, cc[CD::label(u, 1) * M + CD::label(v, 1)]);
            } else {
                t0 = ce0 - po + kb[CD::label(v, 0)];
                t1 = ce1 - po + kb[CD::label(v, 1)];
            }
            r[a] = __viaddmin_u16x2(po, bm[CD::label(v, 0)], ce0);
            r[b] = __viaddmin_u16x2(po, bm[CD::label(v, 1)], ce1);
            uint32_t p;
            asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(p) : "r"(t0), "r"(t1));   // bytes: (A,m) (B,m) (A,m+S/2) (B,m+S/2), 0x00 / 0xFF
            dw[m >> 3] |= p & (0x01010101u << (m & 7));
        }
    });
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

constexpr int kTrace2Chunk = 16;

template <class CD, int MODE>
__global__ void __launch_bounds__(kRegsThreads, kRegs2MinBlocks<CD>)
vit_regs2_kernel(const void* __restrict__ in, uint8_t* __restrict__ out, int32_t* __restrict__ path_metric,
                 uint32_t* __restrict__ surv, uint32_t nframes, uint32_t T, uint32_t out_bytes, uint32_t tile_bytes,
                 uint32_t stride_steps, uint32_t init_all_from) {
    constexpr int S = CD::S, B = CD::B, M = CD::M, NW = Surv2<S>::NW;
    using SV = typename SurvVec<NW>::type;
    extern __shared__ __align__(16) unsigned char s_raw[];      // output staging: [warp][64 rows][out_bytes], tile_bytes per warp
    constexpr bool kLut = VDEC_HARD_LUT && MODE == kModeHard;
    __shared__ __align__(16) uint32_t s_lut[kLut ? kHardLutWords<CD::N> : 4];
    if constexpr (kLut) hard_lut_init<CD::N>(s_lut);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gt = blockIdx.x * kRegsThreads + tid;        // thread <-> frames 2*gt, 2*gt+1
    const uint32_t gw = gt >> 5;
    const uint32_t fA = 2 * gt, fB = fA + 1;
    const bool liveA = fA < nframes, liveB = fB < nframes;
    // rows of the two frames; a dead frame shadows the last live one and stores nothing
    using Elem = std::conditional_t<MODE == kModeHard, uint8_t, uint32_t>;      // one step: 1 symbol byte / 4 metric bytes
    const Elem* rowA = static_cast<const Elem*>(in) + (size_t)min(fA, nframes - 1) * stride_steps;
    const Elem* rowB = static_cast<const Elem*>(in) + (size_t)min(fB, nframes - 1) * stride_steps;
    auto fetch = [&](uint32_t t) { return make_uint2((uint32_t)__ldg(rowA + t), (uint32_t)__ldg(rowB + t)); };
    // the loads above run one step ahead of the ACS, which covers an L1 hit but not the first touch of a 32-byte
    // sector (every 32nd step hard, every 8th quantised): an L1 prefetch kPfSteps further on brings the sector in
    // while the current one is consumed (no register, two LSU slots per B steps).  Measured on the 64-state kernels
    // with true-parity labels (hard 141.4 -> 146.0, quantised 131.6 -> 139.0 Gbit/s; with 168 registers and 3 CTAs per SM,
    // i.e. without their spills, 148.6 and 144.9; rate 1/3 with reference labels 141.8 -> 148.5); with the reference's labels the
    // kernel sits at its instruction-mix limit either way (+-1 %), and on the 4-state codebook trellis, where a
    // B-step group is a few dozen instructions, the two extra LSU slots cost a third (659 -> 430): kRegs2Wide only.
    constexpr uint32_t kPfSteps = kRegs2Wide<CD> ? VDEC_R2_PF_BYTES / sizeof(Elem) : 0;
    auto prefetch_rows = [&](uint32_t t) {
        if constexpr (kPfSteps != 0) {
            const uint32_t tp = min(t + kPfSteps, T - 1);
            asm volatile("prefetch.global.L1 [%0];" :: "l"(rowA + tp));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(rowB + tp));
        }
    };
    SV* sv = reinterpret_cast<SV*>(surv) + (size_t)gw * T * 32 + lane;

    uint32_t r[S];
    r[0] = 0;                                                   // :225
    const uint32_t rest = (fA >= init_all_from ? 0u : (kBig2 & 0xFFFFu)) | (fB >= init_all_from ? 0u : (kBig2 & 0xFFFF0000u));
#pragma unroll
    for (int s = 1; s < S; ++s) r[s] = rest;                    // unreachable (:226-228) unless the start is unknown

    uint32_t dw[NW], bm[M];
    uint2 xn = fetch(0);
    auto store_dw = [&](uint32_t t) {
        SV v;
        if constexpr (NW == 1) v = dw[0];
        else if constexpr (NW == 2) v = make_uint2(dw[0], dw[1]);
        else v = make_uint4(dw[0], dw[1], dw[2], dw[3]);
        if (liveA) sv[(size_t)t * 32] = v;
    };

    uint32_t t = 0;
    for (; t + B <= T; t += B) {
        prefetch_rows(t);
        static_for<0, B>([&](auto phc) {
            constexpr int PH = decltype(phc)::value;
            const uint2 x = xn;
            xn = fetch(min(t + PH + 1, T - 1));                                // one step of lookahead
            if constexpr (kLut) hard_lut_bm<CD::N>(s_lut, x, bm);
            else packed_bm<M, MODE>(x, bm);
            acs_step2<CD, PH>(r, bm, dw);
            store_dw(t + PH);
        });
    }
    const uint32_t rem = T - t;
    static_for<0, B - 1>([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (PH < (int)rem) {
            const uint2 x = xn;
            xn = fetch(min(t + PH + 1, T - 1));                                // one step of lookahead
            if constexpr (kLut) hard_lut_bm<CD::N>(s_lut, x, bm);
            else packed_bm<M, MODE>(x, bm);
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

    // traceback of both frames (:84-93); survivor vectors fetched a chunk ahead.  As in vdec_regs.cuh the
    // state of each frame is an unmasked shift register (cur = h & (S-1)) whose bits B-1.. are the decoded
    // bits, emitted 16 at a time into the warp's tile (64 rows at their final out_bytes pitch).
    unsigned char* tile = s_raw + (size_t)warp * tile_bytes;
    if (liveA) {
        unsigned char* rowA = tile + (size_t)(2 * lane) * out_bytes;
        unsigned char* rowB = rowA + out_bytes;
        SV wc[kTrace2Chunk], wn[kTrace2Chunk];
        int base = ((int)T - 1) & ~(kTrace2Chunk - 1);
#pragma unroll
        for (int j = 0; j < kTrace2Chunk; ++j)
            if (base + j < (int)T) wc[j] = __ldcg(sv + (size_t)(base + j) * 32);
        uint32_t hA = curA, hB = curB;
        auto pick = [&](const SV& v, uint32_t h, uint32_t fr) -> uint32_t {
            const uint32_t cur = h & (S - 1), m = cur & (S / 2 - 1);
            const uint32_t word = surv_word<NW>(v, m >> 3);
            return (word >> (8 * (2 * (cur >> (B - 1)) + fr) + (m & 7))) & 1u;
        };
        auto chunk = [&](auto guard) {
            constexpr bool GUARD = decltype(guard)::value;
#pragma unroll
            for (int j = kTrace2Chunk - 1; j >= 0; --j) {
                const int tt = base + j;
                if (GUARD && tt >= (int)T) continue;
                if ((j & 15) == 0) {
                    const uint32_t ra = __brev(hA >> (B - 1)), rb = __brev(hB >> (B - 1));
                    rowA[tt >> 3] = (unsigned char)(ra >> 24);
                    rowB[tt >> 3] = (unsigned char)(rb >> 24);
                    if (!GUARD || tt + 8 < (int)T) {
                        rowA[(tt >> 3) + 1] = (unsigned char)(ra >> 16);
                        rowB[(tt >> 3) + 1] = (unsigned char)(rb >> 16);
                    }
                }
                hA = (hA << 1) | pick(wc[j], hA, 0);
                hB = (hB << 1) | pick(wc[j], hB, 1);
            }
        };
        auto prefetch = [&]() {
            if (base >= kTrace2Chunk) {
#pragma unroll
                for (int j = 0; j < kTrace2Chunk; ++j) wn[j] = __ldcg(sv + (size_t)(base - kTrace2Chunk + j) * 32);
            }
        };
        prefetch();
        chunk(std::true_type{});
        while (base >= kTrace2Chunk) {
#pragma unroll
            for (int j = 0; j < kTrace2Chunk; ++j) wc[j] = wn[j];
            base -= kTrace2Chunk;
            prefetch();
            chunk(std::false_type{});
        }
    }
    __syncwarp();

    // the warp's 64 output rows are contiguous in global memory
    const uint32_t wbase = 2 * (blockIdx.x * kRegsThreads + (tid & ~31u));
    if (wbase < nframes) {
        const uint32_t rows = min(64u, nframes - wbase);
        const uint32_t total = rows * out_bytes;
        uint8_t* dst = out + (size_t)wbase * out_bytes;
        uint32_t done = 0;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            const uint32_t n16 = total >> 4;
            for (uint32_t i = lane; i < n16; i += 32)
                reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(tile)[i];
            done = n16 << 4;
        }
        for (uint32_t i = done + lane; i < total; i += 32) dst[i] = tile[i];
    }
}

#ifndef __CUDACC_RTC__
struct Regs2Entry {
    const char* name;
    uint32_t K, n, mode;
    uint8_t labels[128];
    int (*launch)(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch, cudaStream_t st);
    size_t (*scratch_bytes)(const LaunchGeom& g);
};
const Regs2Entry* regs2_find(uint32_t K, uint32_t n, uint32_t mode, const uint8_t* labels);

inline size_t regs2_scratch_bytes_nw(const LaunchGeom& g, uint32_t nw) {
    const size_t warps = ((size_t)(g.nframes + 1) / 2 + 31) / 32;
    return warps * g.T * 32 * nw * sizeof(uint32_t);                                     // survivors
}
inline uint32_t regs2_nw(uint32_t K) { return ((1u << (K - 1)) + 15) / 16; }

template <class CD>
size_t regs2_scratch_bytes(const LaunchGeom& g) { return regs2_scratch_bytes_nw(g, Surv2<CD::S>::NW); }

// `kern` is a __global__ symbol of this library or a cudaKernel_t of a run-time compiled module (vdec_jit.cu)
inline int regs2_launch_kernel(const void* kern, const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric,
                               uint32_t* scratch, cudaStream_t st) {
    const size_t npairs = ((size_t)g.nframes + 1) / 2;
    uint32_t tile_bytes = (64 * g.out_bytes + 15) & ~15u;
    const size_t smem = (size_t)(kRegsThreads / 32) * tile_bytes;
    if (smem > 48 * 1024) VDEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (uint32_t)((npairs + kRegsThreads - 1) / kRegsThreads);
    uint32_t nframes = g.nframes, T = g.T, out_bytes = g.out_bytes, stride = g.stride_steps, init_all_from = g.init_all_from;
    void* args[] = {&in, &out, &path_metric, &scratch, &nframes, &T, &out_bytes, &tile_bytes, &stride, &init_all_from};
    VDEC_CUDA(cudaLaunchKernel(kern, dim3(blocks), dim3(kRegsThreads), args, smem, st));
    return 0;
}

template <class CD, int MODE>
int regs2_launch(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch, cudaStream_t st) {
    return regs2_launch_kernel(reinterpret_cast<const void*>(&vit_regs2_kernel<CD, MODE>), g, in, out, path_metric, scratch, st);
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
#endif  // !__CUDACC_RTC__

}  // namespace vd
