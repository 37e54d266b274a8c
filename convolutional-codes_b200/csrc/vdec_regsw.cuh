// vdec_regsw.cuh — register Viterbi kernel for 128 .. 2048 states (K = 8 .. 12): W = 2^(K-7) LANES PER FRAME.
//
// Extends the thread-per-frame scheme of vdec_regs.cuh (AWGN-channel/viterbi-decoder.c:108-141, 71-90) to
// trellises that do not fit one thread's registers.  A state's POSITION after p steps is rotl(s, p) over the
// B = K-1 state bits (in-place butterflies, as before); the low 6 position bits select a register, the high
// WB = B-6 bits select the lane q of the frame's lane group.  The butterfly partner bit is position bit
// (p mod B):
//   * while it is a register bit (6 of every B steps) the step is purely local — the same 2 add / compare /
//     min / decision recipe as the K <= 7 kernel, no communication;
//   * while it is a lane bit (WB of every B steps) partners sit in lanes q and q ^ (1 << tb): one
//     __shfl_xor per register fetches the partner's metric, the lane holding the even state produces the
//     input-0 successor, the other lane the input-1 successor, both in place.
// Labels: the code must be LINEAR (true parity), so label(state, input) = XOR of per-bit contributions; the
// register part is a compile-time constant per (register, phase), the lane part a per-lane runtime constant
// per phase that is applied by permuting the step's M branch metrics (bm[e ^ x]) as they are read from
// shared memory.  (The reference's de-facto labelling is non-linear for polynomials with a D^3 / D^11 tap and
// stays on the table-driven kernels for K >= 8 — the reference itself cannot run K >= 9 at all.)
//
// Survivors: bit r of lane q's 64-bit word at step t is the decision of the state at POSITION (q, r) after
// step t; layout [warp][step][lane] uint2, coalesced.  Traceback: all W lanes of a group walk the same path,
// each prefetching its own words a chunk ahead; the word holding the current state's position is fetched from
// its owner lane with one __shfl.  Tie rule, final lowest-index argmin, output format: as the reference.
#pragma once
#include "vdec_regs.cuh"

namespace vd {

template <int K_, uint64_t P0_, uint64_t P1_>
struct WideDef {
    static constexpr int K = K_, N = 2, M = 4, B = K_ - 1, S = 1 << (K_ - 1), WB = B - 6, W = 1 << WB;
    static_assert(K_ >= 8 && K_ <= 12, "lane-group kernel covers K = 8 .. 12");
    // label contribution of state bit k (0 = oldest register bit) and of the input bit, P0 in the MSB
    VDEC_HD static constexpr uint32_t gbit(int k) {
        return (uint32_t)((((P0_ >> (64 - K_ + k)) & 1ull) << 1) | ((P1_ >> (64 - K_ + k)) & 1ull));
    }
    VDEC_HD static constexpr uint32_t gin() { return (uint32_t)((((P0_ >> 63) & 1ull) << 1) | ((P1_ >> 63) & 1ull)); }
    VDEC_HD static constexpr uint32_t lbits(uint32_t s) {          // label of logical state s, input 0
        uint32_t x = 0;
        for (int k = 0; k < B; ++k) if ((s >> k) & 1u) x ^= gbit(k);
        return x;
    }
    // register part: logical bits contributed by register index r at phase ph
    VDEC_HD static constexpr uint32_t lr(uint32_t r, int ph) { return lbits(rotr_bits(r, ph, B)); }
    VDEC_HD static constexpr uint32_t label(uint32_t s, uint32_t i) { return lbits(s) ^ (i ? gin() : 0u); }
};

template <class WD>
struct WideRing {       // input ring per warp: [slot][frame-of-warp][4 floats]; lane q == 0 of a group feeds it
    static constexpr int G = WD::B, NG = 3, FPW = 32 / WD::W;
    static constexpr uint32_t kSlotBytes = FPW * 16, kGroupBytes = G * kSlotBytes;
    static constexpr size_t kWarpBytes = (size_t)NG * kGroupBytes;
    uint32_t base, T, next_t, woff, roff;
    const float* nextp;
    bool feeder;
    __device__ __forceinline__ void prefetch_next() {
        if (feeder) {
            const uint32_t dst = base + woff;
#pragma unroll
            for (int s = 0; s < G; ++s)
                if (next_t + s < T)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + s * kSlotBytes), "l"(nextp + s * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        nextp += G * 4; next_t += G;
        woff = (woff == (NG - 1) * kGroupBytes) ? 0u : woff + kGroupBytes;
    }
    __device__ __forceinline__ void init(void* warp_smem, uint32_t group, bool is_feeder, const float* row, uint32_t steps) {
        base = (uint32_t)__cvta_generic_to_shared(warp_smem) + group * 16;
        T = steps; nextp = row; next_t = 0; woff = 0; roff = 0; feeder = is_feeder;
        prefetch_next();
        prefetch_next();
    }
    __device__ __forceinline__ void begin_period() {
        __syncwarp();                                  // every lane is done reading the slot about to be refilled
        prefetch_next();
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();                                  // the feeder lane's copies become visible to its group
    }
    __device__ __forceinline__ void end_period() { roff = (roff == (NG - 1) * kGroupBytes) ? 0u : roff + kGroupBytes; }
    // bm[e] = metric[e ^ x] of step `ph` of the current period
    template <int PH>
    __device__ __forceinline__ void get_permuted(uint32_t x, float (&bm)[4]) const {
        const uint32_t a = base + roff + PH * kSlotBytes;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(bm[e]) : "r"(a + 4 * ((uint32_t)e ^ x)));
    }
};

// one step at phase PH for this lane's 64 registers
template <class WD, int PH>
__device__ __forceinline__ void wide_step(float (&r)[64], const WideRing<WD>& ring, const uint32_t (&xq)[WD::B], uint32_t q,
                                          uint32_t (&dw)[2]) {
    constexpr int B = WD::B, PB = PH % B;
    float da[3] = {0.f, 0.f, 0.f};
    if constexpr (PB < 6) {
        // local butterflies (a, a | 1<<PB): every label = compile-time register part ^ this lane's xq[PH]
        float bm[4];
        ring.template get_permuted<PH>(xq[PH], bm);
#pragma unroll
        for (int a = 0; a < 64; ++a) {
            if ((a >> PB) & 1) continue;
            const int b = a | (1 << PB);
            const float pe = r[a], po = r[b];
            const uint32_t la = WD::lr(a, PH), lb = WD::lr(b, PH), gi = WD::gin();
            r[a] = sel_lt(__fadd_rn(pe, bm[la]), __fadd_rn(po, bm[lb]), da[a / kAccBits], acc_weight(a));
            r[b] = sel_lt(__fadd_rn(pe, bm[la ^ gi]), __fadd_rn(po, bm[lb ^ gi]), da[b / kAccBits], acc_weight(b));
        }
    } else {
        // partner lane q ^ (1 << tb); the lane with that bit clear holds the even states and forms the input-0
        // successors, the other the input-1 successors.  pe/po are put in (even, odd) order with two selects.
        constexpr int tb = PB - 6;
        const bool odd = (q >> tb) & 1u;
        const uint32_t xe = xq[PH] ^ (odd ? (WD::gbit(0) ^ WD::gin()) : 0u);     // label offset of the even candidate
        // xq[PH] includes this lane's own bit; the even state's label must not: remove gbit(0) when odd
        float bme[4], bmo[4];
        ring.template get_permuted<PH>(xe, bme);
        ring.template get_permuted<PH>(xe ^ WD::gbit(0), bmo);
#pragma unroll
        for (int a = 0; a < 64; ++a) {
            const float mine = r[a];
            const float other = __shfl_xor_sync(0xffffffffu, mine, 1 << tb);
            const float pe = odd ? other : mine, po = odd ? mine : other;
            const uint32_t la = WD::lr(a, PH);
            r[a] = sel_lt(__fadd_rn(pe, bme[la]), __fadd_rn(po, bmo[la]), da[a / kAccBits], acc_weight(a));
        }
    }
    acc_to_words<64>(da, dw);
}

template <class WD>
__global__ void __launch_bounds__(kRegsThreads, 4)
vit_regsw_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, uint32_t* __restrict__ surv, uint32_t nframes,
                 uint32_t T, uint32_t out_bytes, uint32_t stage_words, uint32_t stride_steps, uint32_t init_all_from) {
    constexpr int B = WD::B, S = WD::S, W = WD::W, FPW = 32 / W;
    using Ring = WideRing<WD>;
    extern __shared__ __align__(16) unsigned char s_raw[];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t q = lane & (W - 1), grp = lane / W;
    const uint32_t gw = blockIdx.x * (kRegsThreads / 32) + warp;          // global warp
    const uint32_t f = gw * FPW + grp;
    const bool live = f < nframes;
    const uint32_t fr = live ? f : nframes - 1;
    const size_t tile_bytes = (size_t)FPW * stage_words * 4;
    const size_t warp_smem = Ring::kWarpBytes > tile_bytes ? Ring::kWarpBytes : tile_bytes;
    unsigned char* my_smem = s_raw + warp * warp_smem;
    uint2* sv = reinterpret_cast<uint2*>(surv) + (size_t)gw * T * 32 + lane;

    // per-phase label offset contributed by this lane's position bits (q << 6)
    uint32_t xq[B];
#pragma unroll
    for (int ph = 0; ph < B; ++ph) {
        const uint32_t s = rotr_bits(q << 6, ph, B);
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < B; ++k) x ^= ((s >> k) & 1u) ? WD::gbit(k) : 0u;
        xq[ph] = x;
    }

    Ring ring;
    ring.init(my_smem, grp, q == 0, in + (size_t)fr * stride_steps * 4, T);

    float r[64];
    const float rest = (f >= init_all_from) ? 0.f : __int_as_float(0x7f800000);
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = rest;
    if (q == 0) r[0] = 0.f;                                               // logical state 0 = position (0, 0) at phase 0 (:215)

    uint32_t dw[2];
    uint32_t t = 0;
    for (; t + B <= T; t += B) {
        ring.begin_period();
        static_for<0, B>([&](auto phc) {
            constexpr int PH = decltype(phc)::value;
            wide_step<WD, PH>(r, ring, xq, q, dw);
            if (live) sv[(size_t)(t + PH) * 32] = make_uint2(dw[0], dw[1]);
        });
        ring.end_period();
    }
    const uint32_t rem = T - t;
    if (rem) ring.begin_period();
    static_for<0, B - 1>([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (PH < (int)rem) {
            wide_step<WD, PH>(r, ring, xq, q, dw);
            if (live) sv[(size_t)(t + PH) * 32] = make_uint2(dw[0], dw[1]);
        }
    });
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    // final state: lowest LOGICAL index among the minima (:72-79); logical = rotr(position, rem)
    float best = __int_as_float(0x7f800000); uint32_t bs = 0xFFFFFFFFu;
    static_for<0, B>([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (PH == (int)rem) {
            const uint32_t sq = rotr_bits(q << 6, PH, B);
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const uint32_t s = sq | rotr_bits((uint32_t)i, PH, B);
                const float v = r[i];
                if (v < best || (v == best && s < bs)) { best = v; bs = s; }
            }
        }
    });
#pragma unroll
    for (int off = 1; off < W; off <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const uint32_t os = __shfl_xor_sync(0xffffffffu, bs, off);
        if (ov < best || (ov == best && os < bs)) { best = ov; bs = os; }
    }
    uint32_t cur = (bs == 0xFFFFFFFFu) ? 0u : bs;
    __syncwarp();

    // traceback: every lane of the group follows the same path; the survivor word of the state's position
    // (lane q', bit r') comes from its owner lane by shuffle.  Words are prefetched a chunk ahead.
    uint32_t* stage = reinterpret_cast<uint32_t*>(my_smem) + (size_t)grp * stage_words;
    {
        uint2 wc[kTraceChunk], wn[kTraceChunk];
        int base = ((int)T - 1) & ~(kTraceChunk - 1);
#pragma unroll
        for (int j = 0; j < kTraceChunk; ++j) {
            wc[j] = make_uint2(0u, 0u);
            if (live && base + j < (int)T) wc[j] = __ldcg(sv + (size_t)(base + j) * 32);
        }
        uint32_t acc = 0;
        uint32_t ph1 = (uint32_t)T % B;                                   // layout phase after step T-1
        for (; base >= 0; base -= kTraceChunk) {
            if (base >= kTraceChunk) {
#pragma unroll
                for (int j = 0; j < kTraceChunk; ++j) {
                    wn[j] = make_uint2(0u, 0u);
                    if (live) wn[j] = __ldcg(sv + (size_t)(base - kTraceChunk + j) * 32);
                }
            }
#pragma unroll
            for (int j = kTraceChunk - 1; j >= 0; --j) {
                const int tt = base + j;
                if (tt < (int)T) {
                    const uint32_t pos = ph1 == 0 ? cur : (((cur << ph1) | (cur >> (B - ph1))) & (S - 1));
                    const uint32_t mine = (pos & 32u) ? wc[j].y : wc[j].x;
                    const uint32_t word = __shfl_sync(0xffffffffu, mine, (lane & ~(uint32_t)(W - 1)) | (pos >> 6));
                    const uint32_t d = (word >> (pos & 31u)) & 1u;
                    acc |= (cur >> (B - 1)) << (31 - (tt & 31));                               // :88
                    cur = ((cur << 1) & (S - 1)) | d;                                           // :86
                    ph1 = ph1 == 0 ? B - 1 : ph1 - 1;
                    if ((tt & 31) == 0) { if (q == 0) stage[tt >> 5] = __byte_perm(acc, 0, 0x0123); acc = 0; }
                }
            }
#pragma unroll
            for (int j = 0; j < kTraceChunk; ++j) wc[j] = wn[j];
        }
    }
    __syncwarp();

    const uint32_t wbase = gw * FPW;                                      // first frame of this warp
    if (wbase < nframes) {
        const uint32_t rows = min((uint32_t)FPW, nframes - wbase);
        const uint32_t total = rows * out_bytes;
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(my_smem);
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

template <class WD>
int regsw_launch(const LaunchGeom& g, const void* in, uint8_t* out, int32_t*, uint32_t* surv, cudaStream_t st) {
    constexpr int FPW = 32 / WD::W;
    uint32_t stage_words = (g.out_bytes + 3) / 4;
    if ((stage_words & 1) == 0) ++stage_words;
    size_t warp_smem = (size_t)FPW * stage_words * 4;
    if (WideRing<WD>::kWarpBytes > warp_smem) warp_smem = WideRing<WD>::kWarpBytes;
    const size_t smem = (size_t)(kRegsThreads / 32) * warp_smem;
    auto kern = vit_regsw_kernel<WD>;
    if (smem > 48 * 1024) VDEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t frames_per_block = (kRegsThreads / 32) * FPW;
    const uint32_t blocks = (g.nframes + frames_per_block - 1) / frames_per_block;
    kern<<<blocks, kRegsThreads, smem, st>>>(static_cast<const float*>(in), out, surv, g.nframes, g.T, g.out_bytes, stage_words,
                                              g.stride_steps, g.init_all_from);
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

template <class WD>
size_t regsw_surv_bytes(const LaunchGeom& g) {
    constexpr int FPW = 32 / WD::W;
    return (size_t)((g.nframes + FPW - 1) / FPW) * g.T * 32 * sizeof(uint2);
}

struct WideEntry {
    const char* name;
    uint32_t K, n;
    const uint8_t* labels;                     // [S][2]
    int (*launch)(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* surv, cudaStream_t st);
    size_t (*surv_bytes)(const LaunchGeom& g);
};

template <class WD>
const uint8_t* wide_labels() {
    static uint8_t store[2 * WD::S];
    for (int s = 0; s < WD::S; ++s) { store[2 * s] = (uint8_t)WD::label(s, 0); store[2 * s + 1] = (uint8_t)WD::label(s, 1); }
    return store;
}

template <class WD>
WideEntry make_wide_entry(const char* name) {
    return WideEntry{name, (uint32_t)WD::K, (uint32_t)WD::N, wide_labels<WD>(), &regsw_launch<WD>, &regsw_surv_bytes<WD>};
}

}  // namespace vd
