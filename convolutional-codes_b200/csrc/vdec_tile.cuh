// vdec_tile.cuh — register-tile Viterbi kernel for large trellises (K = 8 .. 15, 128 .. 16384 states).
//
// Replaces receive_symbol()/traceback() of AWGN-channel/viterbi-decoder.c:108-141, 71-90 for the codes of
// BASELINE config 4 (K = 9 .. 15, survivor memory in HBM; the reference itself hangs for K >= 9, SURVEY finding 2,
// so parity is checked against the widened CPU restatement in tests/).
//
// Forward kernel: 2^(K-7) THREADS PER FRAME, 64 path metrics per thread, all in registers.
//   * In-place butterflies as in vdec_regs.cuh: after p steps logical state s sits at POSITION rotl(s, p) over
//     the B = K-1 state bits, and step p pairs positions that differ in bit p mod B.
//   * Time is cut into WINDOWS of 6 steps.  During window w a thread's register index is position bits
//     6w .. 6w+5 (mod B) and its thread index the remaining B-6 bits, so all 6 pairing bits of the window are
//     register bits: every step is the purely local FADD2 / FSETP / FMNMX / predicated-FADD recipe of the
//     K = 7 kernel (acs_step_f2), no shuffles.  Writing sigma for the position rotated so that the window's
//     bits come first (sigma bit i = position bit 6w+i), sigma bit i holds logical state bit (i-k) mod B at the
//     window's k-th step — independent of w, so ONE 6-step body serves every window.
//   * Between windows the tile is transposed through shared memory: sigma' = rotr(sigma, 6).  64 scalar STS and
//     16 LDS.128 per thread per 6 steps (= 384 state updates), both bank-conflict free through an XOR-swizzled
//     address map phi() that is verified at compile time.  Two group barriers per window (warp-level for
//     K <= 12, named barriers for K = 13, 14, __syncthreads-equivalent for K = 15).
//   * Labels: the code must be LINEAR, label = (register part, compile time) ^ (thread part, 2 bits per step
//     position k, applied by permuting the step's 4 branch metrics as they are read from the input ring).
//   * Decisions: 64 bits per thread per step, stored coalesced to HBM as [frame][step][thread] uint2
//     (K = 15: 2 KB per step and frame); the final state (lowest-index argmin, :72-79) goes to final_state[].
// Traceback kernel: ONE THREAD PER FRAME.  Inside a window all lookups hit the same thread's words (the thread
// index is made of position bits outside the window), so the up to 6 survivor words of a window are fetched
// together and the serial chain sees one HBM latency per window, not per step.
#pragma once
#include "vdec_regs.cuh"

namespace vd {

constexpr int kTileThreads = 256;

VDEC_HD constexpr int ilog2(int v) { int r = 0; while ((1 << r) < v) ++r; return r; }

#ifndef VDEC_TILE_VB
#define VDEC_TILE_VB 2
#endif
// Survivor layout.  Step-major [frame][step][thread] lets the frame's threads store one contiguous record per step, but
// the traceback needs ONE thread's word per step and fetches a 32-byte sector for it.  Thread-major
// [frame][thread][step] makes a window's words 48 contiguous bytes for the traceback; the forward stores become 8-byte
// pieces that L2 merges into sectors.  Measured on the same box (round-1 A/B builds),
// thread-major against step-major: K = 8 -5 %, K = 9 +-0, K = 10 +2.5 %, K = 11 +5.8 %, K = 12 +1.2 %, K = 13 +1 %, K = 14 +1.7 %, K = 15 -1.5 %; hence thread-major
// for 8 .. 128 threads per frame (K = 10 .. 14).
#ifndef VDEC_TILE_TM_MIN
#define VDEC_TILE_TM_MIN 8
#endif
#ifndef VDEC_TILE_TM_MAX
#define VDEC_TILE_TM_MAX 128
#endif
VDEC_HD constexpr bool tile_thread_major(uint32_t tpf) { return tpf >= VDEC_TILE_TM_MIN && tpf <= VDEC_TILE_TM_MAX; }
// Window-major, staged: [warp of the grid][window][lane][6 steps], 48 bytes per lane and window.  The lanes park the six
// words of a window in their warp's (idle) exchange region and the warp writes the 1536-byte record with three fully
// coalesced 16-byte stores per lane; the traceback reads a window's 48 bytes as three 16-byte loads out of two sectors
// instead of six.  Needs a warp-private exchange region, i.e. at most 32 threads per frame (K <= 12).
// Measured against the layouts above (profiles/r02z_ab_tile_window_major.txt): K = 8 +3.3 %, K = 9 +3.5 %, K = 10 +2.6 %,
// K = 11 +2.2 %, K = 12 +0.7 %; K = 9 with the de-facto labels, whose windows already expand their metrics through
// shared memory, -0.9 % — hence true-parity labels only.
#ifndef VDEC_TILE_WM_MIN
#define VDEC_TILE_WM_MIN 1
#endif
#ifndef VDEC_TILE_WM_MAX
#define VDEC_TILE_WM_MAX 32
#endif
enum { kSurvStepMajor = 0, kSurvThreadMajor = 1, kSurvWindowMajor = 2 };
VDEC_HD constexpr int tile_surv_layout(uint32_t tpf, bool ref) {
    return (!ref && tpf <= 32 && tpf >= VDEC_TILE_WM_MIN && tpf <= VDEC_TILE_WM_MAX) ? kSurvWindowMajor
           : tile_thread_major(tpf) ? kSurvThreadMajor : kSurvStepMajor;
}
constexpr uint32_t kWmRecordBytes = 32 * 6 * 8;     // one warp, one window

// N_ = 3 (rate 1/3, 8 metrics per step, third polynomial P2_): true-parity labels only — the de-facto labelling would
// need a 64-letter extended alphabet.
template <int K_, uint64_t P0_, uint64_t P1_, bool REF_ = false, int VB_ = (K_ == 8 ? 1 : VDEC_TILE_VB), int N_ = 2, uint64_t P2_ = 0>
struct TileDef {
    static constexpr bool REF = REF_;                        // the reference's de-facto (non-linear) parity, vdec_labels.h
    static constexpr int K = K_, N = N_, M = 1 << N_, B = K_ - 1, S = 1 << (K_ - 1), TB = B - 6, TPF = 1 << TB;
    static constexpr int ML = REF_ ? 16 : M;                 // size of the LINEAR label alphabet the kernel works in
    static constexpr int VB = VB_, VEC = 1 << VB_;           // registers a reader fetches per LDS (4: LDS.128)
    static constexpr int kStepBytes = M * 4;                 // input bytes per trellis step
    static_assert(N_ == 2 || (N_ == 3 && !REF_), "rate 1/2 (both label modes) or rate 1/3 with true-parity labels");
    static_assert(K_ >= 8 && K_ <= 15, "register-tile kernel covers K = 8 .. 15");
    static_assert(K_ > 8 || VB_ <= 1, "K = 8: the reader's second register bit is a register bit of the writer too, so at most 2 registers per LDS");
    static constexpr int FPC = kTileThreads / TPF;            // frames per CTA
    static constexpr int kLayout = tile_surv_layout(TPF, REF_);           // survivor layout, see VDEC_TILE_TM_MIN / _MAX, VDEC_TILE_WM_MIN / _MAX
    static constexpr bool kThreadMajor = kLayout == kSurvThreadMajor;
    static constexpr bool kWindowMajor = kLayout == kSurvWindowMajor;
    static constexpr bool kSign = K_ >= 9;                    // sign-based decisions (vdec_regs.cuh kDecSign): K = 8 loses 7 % with them
    static constexpr int RT = TPF > 32 ? TPF : 32;            // threads sharing one exchange region
    static constexpr int RTB = ilog2(RT);
    static constexpr int E = 6 + RTB;                         // address bits of a region (RT * 64 floats)
    static constexpr int NREG = kTileThreads / RT;            // exchange regions per CTA

    // label contribution of logical state bit k (0 = oldest register bit) and of the input bit, P0 in the MSB
    VDEC_HD static constexpr uint32_t taps(int pos) {                           // bit `pos` of every polynomial, P0 first
        uint32_t x = (uint32_t)((((P0_ >> pos) & 1ull) << 1) | ((P1_ >> pos) & 1ull));
        if (N_ == 3) x = (x << 1) | (uint32_t)((P2_ >> pos) & 1ull);
        return x;
    }
    VDEC_HD static constexpr uint32_t gbit(int k) { return taps(64 - K_ + k); }
    VDEC_HD static constexpr uint32_t gin() { return taps(63); }
    // De-facto labels (vdec_labels.h: parity_bit): bit j of the label is forced to 0 when z_j = XOR of the tapped
    // register bits at positions 4 mod 8 of the 64-bit shift register is set.  z is linear in the state too, so
    // the kernel carries the 4-bit LINEAR extended label el = parity | z << 2 and maps it to the metric index
    // (el & 3) & ~(el >> 2) only where the branch metrics are fetched.
    VDEC_HD static constexpr uint32_t zbit(int k) {
        const int pos = 64 - K_ + k;
        if (!REF_ || (pos & 7) != 4) return 0u;
        return gbit(k);
    }
    VDEC_HD static constexpr uint32_t ebit(int k) { return gbit(k) | (zbit(k) << 2); }
    VDEC_HD static constexpr uint32_t metric_index(uint32_t el) { return REF_ ? ((el & 3u) & ~(el >> 2)) : el; }
    VDEC_HD static constexpr uint32_t label(uint32_t s, uint32_t i) {          // the n-bit label the decoder compares against
        uint32_t x = i ? gin() : 0u;
        for (int k = 0; k < B; ++k) if ((s >> k) & 1u) x ^= ebit(k);
        return metric_index(x);
    }

    // ---- exchange address map phi: extended sigma (bits 0..5 register, 6.. region-thread index) -> float index.
    // Writer lanes are extended bits 6..10; the reader's lane bit m is extended bit rid(m).  Address bits 0,1 are
    // the reader's register bits 0,1 (sigma 6,7) so that a reader fetches 4 registers with one LDS.128; address
    // bits 2..4 pair the reader's lane bits 0..2 with the writer's lane bits 2..4 by XOR; every other extended
    // bit keeps an address bit of its own.
    VDEC_HD static constexpr int rid(int m) { return m < TB ? (12 + m) % B : 6 + m; }
    VDEC_HD static constexpr bool is_pivot(int e) {
        if (e >= 6 && e < 6 + VB) return true;
        for (int j = 0; j < 5 - VB; ++j) if (e == rid(j)) return true;
        return false;
    }
    VDEC_HD static constexpr uint32_t phi(int e) {
        uint32_t m = 0;
        if (e >= 6 && e < 6 + VB) m |= 1u << (e - 6);
        for (int j = 0; j < 5 - VB; ++j) {
            if (e == rid(j)) m |= (1u << VB) << j;
            else if (e == 6 + VB + j) m |= (1u << VB) << j;
        }
        if (!is_pivot(e)) {
            int rank = 0;
            for (int q = 0; q < e; ++q) if (!is_pivot(q)) ++rank;
            m |= 1u << (5 + rank);
        }
        return m;
    }
    VDEC_HD static constexpr uint32_t cw(uint32_t a) {                 // writer: register a -> address part
        uint32_t x = 0;
        for (int i = 0; i < 6; ++i) if ((a >> i) & 1u) x ^= phi(i);
        return x;
    }
    VDEC_HD static constexpr uint32_t cr(uint32_t a) {                 // reader: new register a -> address part
        uint32_t x = 0;
        for (int i = 0; i < 6; ++i) if ((a >> i) & 1u) x ^= phi((6 + i) % B);
        return x;
    }
    VDEC_HD static constexpr uint32_t wmask() { uint32_t x = 0; for (int m = 0; m < RTB; ++m) x |= phi(6 + m); return x; }
    VDEC_HD static constexpr uint32_t rmask() { uint32_t x = 0; for (int m = 0; m < RTB; ++m) x |= phi(rid(m)); return x; }
    VDEC_HD static constexpr uint32_t wthr(uint32_t thr) { uint32_t x = 0; for (int m = 0; m < RTB; ++m) if ((thr >> m) & 1u) x ^= phi(6 + m); return x; }
    VDEC_HD static constexpr uint32_t rthr(uint32_t thr) { uint32_t x = 0; for (int m = 0; m < RTB; ++m) if ((thr >> m) & 1u) x ^= phi(rid(m)); return x; }

    // compile-time proof obligations of the map
    static constexpr bool phi_bijective() {
        uint32_t rows[E] = {};
        for (int e = 0; e < E; ++e) rows[e] = phi(e);
        int rank = 0;
        for (int bit = 0; bit < E; ++bit) {
            int p = -1;
            for (int r = rank; r < E; ++r) if ((rows[r] >> bit) & 1u) { p = r; break; }
            if (p < 0) return false;
            const uint32_t t = rows[p]; rows[p] = rows[rank]; rows[rank] = t;
            for (int r = 0; r < E; ++r) if (r != rank && ((rows[r] >> bit) & 1u)) rows[r] ^= rows[rank];
            ++rank;
        }
        for (int e = 0; e < E; ++e) if (phi(e) >> E) return false;
        return true;
    }
    static constexpr bool writes_conflict_free() {                      // 32 lanes, one STS.32 each: 32 distinct banks
        uint32_t seen = 0;
        for (uint32_t l = 0; l < 32; ++l) seen |= 1u << (wthr(l) & 31u);
        return seen == 0xFFFFFFFFu;
    }
    static constexpr bool reads_conflict_free() {       // 32/VEC lanes per wavefront group cover distinct VEC-word bank groups
        constexpr uint32_t lanes = 32u / VEC;
        for (uint32_t hi = 0; hi < (uint32_t)VEC; ++hi) {
            uint32_t seen = 0;
            for (uint32_t l = 0; l < lanes; ++l) {
                const uint32_t a = rthr(hi * lanes + l);
                if (a & (VEC - 1u)) return false;
                seen |= 1u << ((a >> VB) & (lanes - 1u));
            }
            if (seen != (lanes == 32u ? 0xFFFFFFFFu : (1u << lanes) - 1u)) return false;
        }
        return true;
    }
    static_assert(phi_bijective(), "exchange address map must be a bijection");
    static_assert(writes_conflict_free(), "exchange writes must be bank-conflict free");
    static_assert(reads_conflict_free(), "exchange reads must be bank-conflict free");
};

// The 64 registers of a thread at the window's k-th step, seen as a K = 7 code: local logical bit c of
// u = rotr6(register, k) is true logical bit c for c <= 5-k and bit B-6+c above (the bits that wrapped).
template <class TD, int KS>
struct TileCode {
    static constexpr int K = 7, N = TD::N, B = 6, S = 64, M = TD::ML, NW = 2;
    static constexpr bool kSignDecisions = TD::kSign;                          // the traceback kernel is told (TileShape::sign)
    VDEC_HD static constexpr uint32_t label(uint32_t u, uint32_t i) {          // register part of the (extended) label
        uint32_t x = i ? TD::gin() : 0u;
        for (int c = 0; c < 6; ++c)
            if ((u >> c) & 1u) x ^= TD::ebit(c <= 5 - KS ? c : TD::B - 6 + c);
        return x;
    }
};

template <int TPF>
__device__ __forceinline__ void tile_group_sync(uint32_t group) {
    if constexpr (TPF <= 32) __syncwarp();
    else if constexpr (TPF == kTileThreads) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(TPF) : "memory");
}

// operand pairs (vm[x ^ xt], vm[y ^ xt]) for the label pairs phase KS needs; slot = shared-window address of the
// step's ML entries (the 4 branch metrics, or their 16-entry expansion by extended label; entries 2^ESH bytes
// apart), aligned to its size
template <class CD, int KS, int ESH = 2>
__device__ __forceinline__ void tile_get_pairs(uint32_t slot, uint32_t xt, uint64_t (&bp)[CD::M * CD::M]) {
    constexpr int M = CD::M;
    const uint32_t bx = slot ^ (xt << ESH);
    static_for<0, M * M>([&](auto ic) {
        constexpr int I = decltype(ic)::value, x = I / M, y = I % M;
        if constexpr (pair_used<CD, KS>(I)) {
            float lo, hi;
            asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(lo) : "r"(bx ^ ((uint32_t)x << ESH)));
            asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(hi) : "r"(bx ^ ((uint32_t)y << ESH)));
            bp[I] = pack2(lo, hi);
        } else {
            bp[I] = 0;
        }
    });
}

constexpr int kTileRingGroups = 3;                                  // windows of input resident per frame
VDEC_HD constexpr int tile_ring_bytes(int n) { return kTileRingGroups * 6 * (4 << n); }     // per frame
constexpr int kTileVmBytes = 6 * 16 * 4;                            // per frame: one window of metrics expanded by extended label

template <class TD>
__global__ void __launch_bounds__(kTileThreads, 2)
vit_tile_forward(const float* __restrict__ in, uint2* __restrict__ surv, uint32_t* __restrict__ final_state,
                 uint32_t nframes, uint32_t T, uint32_t stride_steps, uint32_t init_all_from) {
    constexpr int B = TD::B, S = TD::S, TB = TD::TB, TPF = TD::TPF, FPC = TD::FPC, RT = TD::RT;
    extern __shared__ __align__(1024) unsigned char s_raw[];
    float* xall = reinterpret_cast<float*>(s_raw);                               // [NREG][RT * 64] exchange regions
    constexpr int SB = TD::kStepBytes, GB = 6 * SB, kTileRingBytes = tile_ring_bytes(TD::N);
    unsigned char* rings = s_raw + (size_t)kTileThreads * 64 * 4;               // [FPC][kTileRingBytes]
    // REF: metrics expanded by extended label, [warp or frame][6 steps][16 labels][FPW frames of the warp] floats —
    // frames interleaved innermost so that the lanes of a warp spread over the banks; aligned for XOR addressing
    constexpr int FPW = TPF < 32 ? 32 / TPF : 1, FSH = ilog2(FPW);
    unsigned char* vms = rings + (((size_t)FPC * kTileRingBytes + 511) & ~(size_t)511);
    float* red_v = reinterpret_cast<float*>(vms + (TD::REF ? (size_t)FPC * kTileVmBytes : 0));   // [kTileThreads / 32]
    uint32_t* red_s = reinterpret_cast<uint32_t*>(red_v + kTileThreads / 32);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t tau = tid & (TPF - 1), fic = tid / TPF;
    const uint32_t f = blockIdx.x * FPC + fic;
    const bool live = f < nframes;
    const uint32_t fr = live ? f : nframes - 1;                                 // dead groups shadow the last frame
    const uint32_t thr = tid & (RT - 1);
    float* xbuf = xall + (size_t)(tid / RT) * RT * 64;
    const uint32_t wbase = TD::wthr(thr), rbase = TD::rthr(thr);

    // thread part of the (extended) labels at step position k: tau bit j is sigma bit 6+j = logical bit 6+j-k
    uint32_t xt[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < TB; ++j) x ^= ((tau >> j) & 1u) ? TD::ebit(6 + j - k) : 0u;
        xt[k] = x;
    }

    // input ring of this frame: [group][step][M floats], filled by the frame's thread 0
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(rings + (size_t)fic * kTileRingBytes);
    [[maybe_unused]] const uint32_t fiw = fic & (FPW - 1);                       // frame within its vm group
    [[maybe_unused]] float* vm = reinterpret_cast<float*>(vms + (size_t)(fic >> FSH) * FPW * kTileVmBytes) + fiw;
    [[maybe_unused]] const uint32_t vma = (uint32_t)__cvta_generic_to_shared(vm);
    const float* nextp = in + (size_t)fr * stride_steps * TD::M;
    uint32_t next_t = 0, woff = 0, roff = 0;
    auto prefetch = [&]() {
        if (tau == 0) {
#pragma unroll
            for (int s = 0; s < 6; ++s)
                if (next_t + s < T) {
#pragma unroll
                    for (int c = 0; c < SB / 16; ++c)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ring + woff + s * SB + c * 16), "l"(nextp + s * TD::M + c * 4) : "memory");
                }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        nextp += 6 * TD::M; next_t += 6;
        woff = (woff == (kTileRingGroups - 1) * GB) ? 0u : woff + GB;
    };
    prefetch();
    prefetch();

    // decoder_reset (:212-222): state 0 (sigma = 0: thread 0, register 0) starts at 0, the rest at +INF
    const float rest = (f >= init_all_from) ? 0.f : __int_as_float(0x7f800000);
    uint64_t rp[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) rp[q] = pack2(rest, rest);
    if (tau == 0) rp[0] = pack2(0.f, rest);

    uint2* sv = TD::kThreadMajor ? surv + ((size_t)f * TPF + tau) * T : surv + (size_t)f * T * TPF + tau;
    // window-major: this warp's records, one per window; the words of the current window wait in the exchange region
    [[maybe_unused]] uint4* wmp = reinterpret_cast<uint4*>(surv) +
                                  (size_t)(blockIdx.x * (kTileThreads / 32) + (tid >> 5)) * ((T + 5) / 6) * (kWmRecordBytes / 16) + lane;
    [[maybe_unused]] uint2* stage2 = reinterpret_cast<uint2*>(xbuf) + lane * 6;
    [[maybe_unused]] const uint4* stage4 = reinterpret_cast<const uint4*>(xbuf) + lane;
    auto flush_window = [&]() {
        if constexpr (TD::kWindowMajor) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 3; ++j) __stcg(wmp + j * 32, stage4[j * 32]);
            wmp += kWmRecordBytes / 16;
            __syncwarp();                                   // the transpose reuses the region
        }
    };
    uint32_t t = 0;
    auto begin_window = [&]() {
        prefetch();
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        tile_group_sync<TPF>(fic);                      // ring data visible; the previous exchange has been read by all
        if constexpr (TD::REF) {                        // expand the window's metrics: vm[step][el] = bm[step][(el & 3) & ~(el >> 2)]
            const float* raw = reinterpret_cast<const float*>(rings + (size_t)fic * kTileRingBytes) + roff / 4;
#pragma unroll
            for (int e = 0; e < (96 + TPF - 1) / TPF; ++e) {
                const uint32_t i = e * TPF + tau;
                if (TPF >= 96 ? tau < 96u : i < 96u) vm[(size_t)i << FSH] = raw[(i >> 4) * 4 + TD::metric_index(i & 15u)];
            }
            tile_group_sync<TPF>(fic);
        }
    };
    auto step = [&](auto kc) {
        constexpr int KS = decltype(kc)::value;
        using CD = TileCode<TD, KS>;
        uint64_t bp[CD::M * CD::M];
        uint32_t dw[2];
        if constexpr (TD::REF) tile_get_pairs<CD, KS, 2 + FSH>(vma + KS * (64 << FSH), xt[KS], bp);
        else tile_get_pairs<CD, KS>(ring + roff + KS * SB, xt[KS], bp);
        acs_step_f2<CD, KS>(rp, bp, dw);
        if constexpr (TD::kWindowMajor) stage2[KS] = make_uint2(dw[0], dw[1]);
        else if (live) sv[TD::kThreadMajor ? (size_t)(t + KS) : (size_t)(t + KS) * TPF] = make_uint2(dw[0], dw[1]);
    };
    // full windows that are followed by another one: straight-line 6 steps, then the transpose
    for (; t + 6 < T; t += 6) {
        begin_window();
        static_for<0, 6>(step);
        flush_window();
        roff = (roff == (kTileRingGroups - 1) * GB) ? 0u : roff + GB;
        // ---- transpose: sigma' = rotr(sigma, 6) -------------------------------------------------------
        static_for<0, 64>([&](auto ac) {
            constexpr uint32_t A = decltype(ac)::value, c = TD::cw(A), cx = c & TD::wmask(), cd = c & ~TD::wmask();
            float lo, hi;
            unpack2(rp[A >> 1], lo, hi);
            xbuf[(wbase ^ cx) + cd] = (A & 1u) ? hi : lo;
        });
        tile_group_sync<TPF>(fic);
        static_for<0, 64 / TD::VEC>([&](auto vc) {
            constexpr uint32_t V = decltype(vc)::value, c = TD::cr(TD::VEC * V), cx = c & TD::rmask(), cd = c & ~TD::rmask();
            const float* src = xbuf + ((rbase ^ cx) + cd);
            if constexpr (TD::VEC == 4) {
                const float4 q = *reinterpret_cast<const float4*>(src);
                rp[2 * V] = pack2(q.x, q.y);
                rp[2 * V + 1] = pack2(q.z, q.w);
            } else if constexpr (TD::VEC == 2) {
                const float2 q = *reinterpret_cast<const float2*>(src);
                rp[V] = pack2(q.x, q.y);
            } else {
                float lo, hi;
                unpack2(rp[V >> 1], lo, hi);
                if (V & 1u) hi = *src; else lo = *src;
                rp[V >> 1] = pack2(lo, hi);
            }
        });
    }
    // last window: 1 .. 6 steps
    const uint32_t n = T - t;
    begin_window();
    static_for<0, 6>([&](auto kc) { if ((int)n > (decltype(kc)::value)) step(kc); });
    flush_window();
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    // ---- final state: lowest LOGICAL index among the minima (:72-79); logical = rotr_B(sigma, k+1) ---------
    float best = __int_as_float(0x7f800000);
    uint32_t bs = 0xFFFFFFFFu;
    static_for<0, 6>([&](auto kc) {
        constexpr int KS = decltype(kc)::value;
        if (KS + 1 == (int)n) {
            const uint32_t x = tau << 6;
            const uint32_t st = ((x >> (KS + 1)) | (x << (B - KS - 1))) & (S - 1);
#pragma unroll
            for (int a = 0; a < 64; ++a) {
                float lo, hi;
                unpack2(rp[a >> 1], lo, hi);
                const float v = (a & 1) ? hi : lo;
                const uint32_t s = st | rotr_bits((uint32_t)a, KS + 1, B);
                if (v < best || (v == best && s < bs)) { best = v; bs = s; }
            }
        }
    });
#pragma unroll
    for (int off = 1; off < (TPF < 32 ? TPF : 32); off <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const uint32_t os = __shfl_xor_sync(0xffffffffu, bs, off);
        if (ov < best || (ov == best && os < bs)) { best = ov; bs = os; }
    }
    if constexpr (TPF > 32) {
        if (lane == 0) { red_v[tid >> 5] = best; red_s[tid >> 5] = bs; }
        __syncthreads();
        if (tau == 0) {
#pragma unroll
            for (int w = 1; w < TPF / 32; ++w) {
                const float ov = red_v[(tid >> 5) + w];
                const uint32_t os = red_s[(tid >> 5) + w];
                if (ov < best || (ov == best && os < bs)) { best = ov; bs = os; }
            }
        }
    }
    if (live && tau == 0) final_state[f] = (bs == 0xFFFFFFFFu) ? 0u : bs;
}

// ONE THREAD PER FRAME.  At the window's k-th step the decision of state `cur` is bit
// (cur & (2^(5-k) - 1)) | (cur >> (B-1-k)) << (5-k) of thread (cur >> (5-k)) & (TPF-1)'s word; the thread index is
// the same for every step of a window, so a window's words are fetched together.  Output as in vdec_regs.cuh:
// the state doubles as a shift register of decoded bits, 16 of which are emitted every 16 steps (:81-89).
// LAYOUT: survivor layout (tile_surv_layout); SIGN: rotated survivor-bit layout (TileDef::kSign)
template <int LAYOUT, bool SIGN>
static __global__ void __launch_bounds__(128)
vit_tile_traceback(const uint2* __restrict__ surv, const uint32_t* __restrict__ final_state, uint8_t* __restrict__ out,
                   uint32_t nframes, uint32_t T, uint32_t out_bytes, uint32_t B, uint32_t TPF) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const uint2* sv = surv + (size_t)f * T * TPF;
    // window-major: the frame's threads are threads (f % FPC) * TPF .. of block f / FPC of the forward grid
    [[maybe_unused]] const uint32_t FPC = kTileThreads / TPF, nwin = (T + 5) / 6;
    [[maybe_unused]] const uint32_t tid0 = (f % FPC) * TPF, blk = f / FPC;
    uint8_t* row = out + (size_t)f * out_bytes;
    const uint32_t S1 = (1u << B) - 1u;
    uint32_t h = final_state[f];
    for (int w0 = (int)((T - 1) / 6) * 6; w0 >= 0; w0 -= 6) {
        const int n = (int)T - w0 < 6 ? (int)T - w0 : 6;
        const uint32_t owner = ((h & S1) >> (6 - n)) & (TPF - 1);
        uint2 wd[6];
        if constexpr (LAYOUT == kSurvWindowMajor) {
            const uint32_t otid = tid0 + owner;
            const uint4* rec = reinterpret_cast<const uint4*>(surv) +
                               ((size_t)(blk * (kTileThreads / 32) + (otid >> 5)) * nwin + (uint32_t)w0 / 6) * (kWmRecordBytes / 16) + (otid & 31u) * 3;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint4 q = __ldcg(rec + j);
                wd[2 * j] = make_uint2(q.x, q.y);
                wd[2 * j + 1] = make_uint2(q.z, q.w);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (k < n) wd[k] = __ldcg(LAYOUT == kSurvThreadMajor ? sv + (size_t)owner * T + (w0 + k) : sv + (size_t)(w0 + k) * TPF + owner);
        }
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
            uint32_t bit = (cur & ((1u << (5 - k)) - 1u)) | ((cur >> (B - 1 - k)) << (5 - k));
            if constexpr (VDEC_DEC_SIGN && SIGN) bit = rotl_bits(bit, dec_rot(k), 6);   // survivor bit layout of acs_step_f2 (vdec_regs.cuh)
            const uint32_t word = (bit & 32u) ? wd[k].y : wd[k].x;
            h = (h << 1) | ((word >> (bit & 31u)) & 1u);                 // predecessor (:86)
        }
    }
}

#ifndef __CUDACC_RTC__
// ---- host side ------------------------------------------------------------------------------------
// Launch shape as a function of (K, label mode) only: kernels compiled at run time (vdec_jit.cu) go through the same launcher.
struct TileShape {
    uint32_t B, TPF, FPC;
    bool ref;
    bool sign;              // survivor bits in the rotated layout of the sign-based decisions (TileDef::kSign)
    uint32_t n;             // coded bits per step: input ring bytes
};
inline TileShape tile_shape(uint32_t K, bool ref, uint32_t n = 2) {
    const uint32_t tpf = 1u << (K - 7);
    return TileShape{K - 1, tpf, (uint32_t)kTileThreads / tpf, ref, K >= 9, n};
}
// survivor words of one launch: per frame T * TPF words, or whole records of every warp of the forward grid (window-major)
inline size_t tile_surv_only_bytes(const LaunchGeom& g, const TileShape& sh) {
    if (tile_surv_layout(sh.TPF, sh.ref) == kSurvWindowMajor) {
        const size_t blocks = (g.nframes + sh.FPC - 1) / sh.FPC;
        return blocks * (kTileThreads / 32) * ((g.T + 5) / 6) * kWmRecordBytes;
    }
    return (size_t)g.nframes * g.T * sh.TPF * sizeof(uint2);
}
inline size_t tile_surv_bytes_shape(const LaunchGeom& g, const TileShape& sh) {
    return tile_surv_only_bytes(g, sh) + (size_t)g.nframes * sizeof(uint32_t);
}

// `kern`: vit_tile_forward<TD> of this library or a cudaKernel_t of a run-time compiled module
inline int tile_launch_kernel(const void* kern, const TileShape& sh, const LaunchGeom& g, const void* in, uint8_t* out,
                              uint32_t* scratch, cudaStream_t st) {
    uint2* surv = reinterpret_cast<uint2*>(scratch);
    uint32_t* final_state = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(scratch) + tile_surv_only_bytes(g, sh));
    const size_t smem = (size_t)kTileThreads * 64 * 4 + (((size_t)sh.FPC * tile_ring_bytes((int)sh.n) + 511) & ~(size_t)511) +
                        (sh.ref ? (size_t)sh.FPC * kTileVmBytes : 0) + (kTileThreads / 32) * 8;
    VDEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (g.nframes + sh.FPC - 1) / sh.FPC;
    const float* inf = static_cast<const float*>(in);
    uint32_t nframes = g.nframes, T = g.T, stride = g.stride_steps, init_all_from = g.init_all_from;
    void* args[] = {&inf, &surv, &final_state, &nframes, &T, &stride, &init_all_from};
    VDEC_CUDA(cudaLaunchKernel(kern, dim3(blocks), dim3(kTileThreads), args, smem, st));
    const uint32_t tb_blocks = (g.nframes + 127) / 128;
    const int layout = tile_surv_layout(sh.TPF, sh.ref);
#define VDEC_TILE_TB(L, SG) vit_tile_traceback<L, SG><<<tb_blocks, 128, 0, st>>>(surv, final_state, out, g.nframes, g.T, g.out_bytes, sh.B, sh.TPF)
    if (layout == kSurvWindowMajor) { if (sh.sign) VDEC_TILE_TB(kSurvWindowMajor, true); else VDEC_TILE_TB(kSurvWindowMajor, false); }
    else if (layout == kSurvThreadMajor) { if (sh.sign) VDEC_TILE_TB(kSurvThreadMajor, true); else VDEC_TILE_TB(kSurvThreadMajor, false); }
    else { if (sh.sign) VDEC_TILE_TB(kSurvStepMajor, true); else VDEC_TILE_TB(kSurvStepMajor, false); }
#undef VDEC_TILE_TB
    VDEC_CUDA(cudaGetLastError());
    return 0;
}

template <class TD>
constexpr TileShape tile_shape_of() { return TileShape{(uint32_t)TD::B, (uint32_t)TD::TPF, (uint32_t)TD::FPC, TD::REF, TD::kSign, (uint32_t)TD::N}; }
template <class TD>
size_t tile_surv_bytes(const LaunchGeom& g) { return tile_surv_bytes_shape(g, tile_shape_of<TD>()); }

template <class TD>
int tile_launch(const LaunchGeom& g, const void* in, uint8_t* out, uint32_t* scratch, cudaStream_t st) {
    return tile_launch_kernel(reinterpret_cast<const void*>(&vit_tile_forward<TD>), tile_shape_of<TD>(), g, in, out, scratch, st);
}

struct TileEntry {
    const char* name;
    uint32_t K, n;
    const uint8_t* labels;                     // [S][2]
    int (*launch)(const LaunchGeom& g, const void* in, uint8_t* out, uint32_t* scratch, cudaStream_t st);
    size_t (*bytes)(const LaunchGeom& g);
};

template <class TD>
const uint8_t* tile_labels() {
    static uint8_t store[2 * TD::S];
    for (int s = 0; s < TD::S; ++s) { store[2 * s] = (uint8_t)TD::label(s, 0); store[2 * s + 1] = (uint8_t)TD::label(s, 1); }
    return store;
}

template <class TD>
TileEntry make_tile_entry(const char* name) {
    return TileEntry{name, (uint32_t)TD::K, (uint32_t)TD::N, tile_labels<TD>(), &tile_launch<TD>, &tile_surv_bytes<TD>};
}

// Labels of the kernel's label model (TileDef::label) for run-time polynomials: a run-time compiled kernel serves a
// request only when this equals the table computed from the caller's polynomials (vdec_labels.h).
inline void tile_model_labels(uint32_t K, uint64_t p0, uint64_t p1, bool ref, uint8_t* labels, uint32_t n = 2, uint64_t p2 = 0) {
    const uint32_t B = K - 1, S = 1u << B;
    auto taps = [&](int pos) {
        uint32_t x = (uint32_t)((((p0 >> pos) & 1ull) << 1) | ((p1 >> pos) & 1ull));
        return n == 3 ? ((x << 1) | (uint32_t)((p2 >> pos) & 1ull)) : x;
    };
    auto ebit = [&](int k) { const int pos = 64 - (int)K + k; return taps(pos) | ((ref && (pos & 7) == 4) ? taps(pos) << 2 : 0u); };
    const uint32_t gin = taps(63);
    for (uint32_t s = 0; s < S; ++s)
        for (uint32_t i = 0; i < 2; ++i) {
            uint32_t x = i ? gin : 0u;
            for (uint32_t k = 0; k < B; ++k) if ((s >> k) & 1u) x ^= ebit((int)k);
            labels[2 * s + i] = (uint8_t)(ref ? ((x & 3u) & ~(x >> 2)) : x);
        }
}
#endif  // !__CUDACC_RTC__

}  // namespace vd
