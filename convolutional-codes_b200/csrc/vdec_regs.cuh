// vdec_regs.cuh — code-specialised Viterbi kernel: ONE THREAD PER FRAME.
//
// Replaces receive_symbol()/traceback() of the reference decoders
// (AWGN-channel/viterbi-decoder.c:71-90,108-141; binary-symmetric-channel/viterbi-decoder.c:75-96,
// 114-150) for codes with S = 2^(K-1) <= 64 states.
//
// Layout of the computation
//   * all S path metrics of a frame live in registers of one thread; a warp decodes 32 frames,
//     so no shuffles, no shared-memory traffic and no divergence in the ACS loop;
//   * the trellis butterfly {2m,2m+1} -> {m, m+S/2} is done IN PLACE: new[m] overwrites the
//     register of old[2m], new[m+S/2] that of old[2m+1].  After p steps logical state s sits in
//     physical register rotl(s, p) (rotation over the K-1 state bits), so unrolling K-1 steps
//     returns to the identity and the loop needs no register moves;
//   * branch labels are compile-time constants (template parameter CodeDef), so `bm[label]`
//     is a register operand: per state update the kernel issues 2 adds, 1 compare, 1 select and
//     1 predicated OR that drops the decision bit into its survivor word;
//   * decisions: 1 bit per state per step (the reference stores a 4-byte struct,
//     viterbi-decoder.c:12-15), written coalesced as [warp][step][lane] words to an HBM scratch;
//   * tie rule = reference: strict '<' with the even predecessor visited first => a tie keeps
//     predecessor 2m, decision bit 0 (viterbi-decoder.c:115-126); final state = lowest-index
//     argmin over all states (:72-79); every one of the T steps allows both inputs;
//   * traceback re-reads the scratch (coalesced), emits all T bits MSB-first (:81-89) through a
//     shared-memory staging tile so the global write of the warp's 32 output rows is contiguous.
//
// Soft metrics are fp32 and every candidate is `pm + bm` exactly as the reference computes it
// (:121), so the result is bit-identical, not approximately equal.  Inputs are assumed finite
// and NaN-free (the demapper guarantees it).
#pragma once
#include <type_traits>
#include <utility>

#include "vdec_internal.h"
#include "vdec_demap.cuh"

// tuning knobs for A/B builds (tools/build_variant.py)
#ifndef VDEC_MINB_REF
#define VDEC_MINB_REF 4
#endif
#ifndef VDEC_MINB_LIN
#define VDEC_MINB_LIN 3
#endif
#ifndef VDEC_SIGN_LIN
#define VDEC_SIGN_LIN 0
#endif
#ifndef VDEC_OPERAND_V4
#define VDEC_OPERAND_V4 1
#endif
// cp.async.ca (allocate in L1) instead of .cg for the 16-byte pieces of the input ring.  A 32-byte sector holds two
// pieces (two consecutive steps at n = 2, the two halves of a step at n = 3) and .cg fetches it from L2 once per
// piece (ncu r02p: 41 KB per frame cross the L2 -> L1 crossbar for 24.7 KB of DRAM reads); with .ca the second
// request hits L1.  Rate 1/3 needs the ring cut to two groups for that: three groups are 72 KB per CTA, 216 KB per SM,
// and leave no L1 to hit in (.ca alone: 75.6 -> 65.4 Gbit/s).  Same-box A/B (profiles/r02t_ab_ring_l1_policy.txt):
// K = 7 reference labels 118.7 - 121.2 -> 123.0 - 123.7, true-parity 98.6 -> 99.6 - 100.4, rate 1/3 75.2 - 76.3 -> 82.2 - 82.4
// (true-parity labels 77.5 - 78.6 -> 87.6 - 88.6).
#ifndef VDEC_RING_CA
#define VDEC_RING_CA 1
#endif
#ifndef VDEC_RING_CA_M8
#define VDEC_RING_CA_M8 1
#endif
#ifndef VDEC_RING_NG_M4
#define VDEC_RING_NG_M4 3        // groups of the rate-1/2 ring
#endif
#ifndef VDEC_RING_NG_M8
#define VDEC_RING_NG_M8 2        // groups of the rate-1/3 ring
#endif
// First-wave stagger (see vit_regs_kernel): SM clocks per trellis step by which the resident CTAs of an SM are spread
// over one CTA lifetime; 0 = off.  Same-box A/B (profiles/r02g_ab_stagger.txt): K = 7 reference labels 119.4 -> 122.0
// Gbit/s, rate 1/3 74.9 -> 76.2, true-parity rate 1/2 99.3 -> 98.5 (stays off there).
#ifndef VDEC_STAGGER
#define VDEC_STAGGER 670
#endif

namespace vd {

template <int K_, int N_, uint64_t P0_, uint64_t P1_, uint64_t P2_, bool REF_>
struct CodeDef {
    static constexpr int K = K_, N = N_, B = K_ - 1, S = 1 << (K_ - 1), M = 1 << N_;
    static constexpr int NW = (S + 31) / 32;
    static constexpr bool REF = REF_;
    // resident CTAs per SM the register budget is cut for.  Measured (tools/bench_configs.py, A/B of
    // builds): labellings with fewer shared candidate sums, and n = 3, spill at 128 registers and run
    // 5 % faster with 168 registers and 3 CTAs; the reference-labelled (171,133) code prefers 4.
    static constexpr int kMinBlocks = (K_ == 7 && !(N_ == 2 && REF_)) ? VDEC_MINB_LIN : VDEC_MINB_REF;
    static constexpr bool kSignDecisions = (N_ == 2 && (REF_ || VDEC_SIGN_LIN));      // survivor bits from packed differences, see kDecSign
    static constexpr int kStagger = (K_ == 7 && (REF_ || N_ == 3)) ? VDEC_STAGGER : 0;  // first-wave stagger, see VDEC_STAGGER
    VDEC_HD static constexpr uint32_t label(uint32_t s, uint32_t i) {
        const uint64_t p[3] = {P0_, P1_, P2_};
        return branch_label(K_, N_, p, REF_, s, i);
    }
};

#ifndef VDEC_REGS_THREADS
#define VDEC_REGS_THREADS 128
#endif
constexpr int kRegsThreads = VDEC_REGS_THREADS;

// static_for<I, N>(f): compile-time loop, defined in vdec_demap.cuh

// ---- metric policies -----------------------------------------------------------------------
struct SoftF32 {
    using T = float;
    using In = float;
    static constexpr bool kPacked = true;     // candidates formed with FADD2 on register pairs
    static constexpr uint32_t kMetric = VDEC_SOFT_F32;
    __device__ static __forceinline__ T init_rest() { return __int_as_float(0x7f800000); }   // +INF, :215-218
    // dst = min(ce, co), decision bit set iff co < ce  (strict: tie -> even predecessor)
    __device__ static __forceinline__ void acs(T& dst, T pe, T po, T be, T bo, uint32_t& dw, uint32_t bit) {
        const float ce = __fadd_rn(pe, be), co = __fadd_rn(po, bo);
        asm("{\n\t.reg .pred p;\n\t"
            "setp.lt.f32 p, %3, %2;\n\t"
            "selp.f32 %0, %3, %2, p;\n\t"
            "@p or.b32 %1, %1, %4;\n\t}"
            : "=f"(dst), "+r"(dw) : "f"(ce), "f"(co), "r"(bit));
    }
    __device__ static __forceinline__ bool less(T a, T b) { return a < b; }
    __device__ static __forceinline__ int32_t to_path_metric(T) { return 0; }
    // branch metrics of one step: the M demapper distances, as loaded
    template <int M>
    __device__ static __forceinline__ void load(const In* __restrict__ p, T (&bm)[M]) {
        if constexpr (M == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); bm[0] = v.x; bm[1] = v.y; }
        else {
#pragma unroll
            for (int i = 0; i < M / 4; ++i) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
                bm[4 * i] = v.x; bm[4 * i + 1] = v.y; bm[4 * i + 2] = v.z; bm[4 * i + 3] = v.w;
            }
        }
    }
    template <int M> VDEC_HD static constexpr size_t step_elems() { return M; }
};

struct HardU16 {
    using T = uint32_t;      // uint16 semantics (clamp at 0xFF00) carried in 32-bit registers
    using In = uint8_t;
    static constexpr bool kPacked = false;
    static constexpr uint32_t kMetric = VDEC_HARD_U16;
    __device__ static __forceinline__ T init_rest() { return 0xFF00u; }                      // :225-228
    __device__ static __forceinline__ void acs(T& dst, T pe, T po, T be, T bo, uint32_t& dw, uint32_t bit) {
        const uint32_t ce = min(pe + be, 0xFF00u), co = min(po + bo, 0xFF00u);                  // :127-130
        const bool d = co < ce;
        dst = d ? co : ce;
        if (d) dw |= bit;
    }
    __device__ static __forceinline__ bool less(T a, T b) { return a < b; }
    __device__ static __forceinline__ int32_t to_path_metric(T v) { return (int32_t)v; }
    // branch metrics of one step: Hamming distance of every possible label to the received symbol (:69-72)
    template <int M>
    __device__ static __forceinline__ void load(const In* __restrict__ p, T (&bm)[M]) {
        const uint32_t r = __ldg(p) & (M - 1);                                                 // :235-237
#pragma unroll
        for (int e = 0; e < M; ++e) bm[e] = __popc(e ^ r);
    }
    template <int M> VDEC_HD static constexpr size_t step_elems() { return 1; }
};

// Soft decisions from RAW I/Q: the input is what the channel hands to the demapper
// (AWGN-channel/main.c:104), float32 [frame][step][2]; the M branch metrics of a step are formed in the feeder
// (IqRing) with the demapper's own arithmetic (vdec_demap.cuh), so the decoder reads 8 B per step instead of 4 M.
struct SoftIq : SoftF32 {
    static constexpr bool kIq = true;
    template <int M> VDEC_HD static constexpr size_t step_elems() { return 2; }
};
template <class MP> constexpr bool kIsIq = false;
template <> constexpr bool kIsIq<SoftIq> = true;

// ---- one trellis step at rotation phase PH ---------------------------------------------------
template <class CD, class MP, int PH>
__device__ __forceinline__ void acs_step(typename MP::T (&r)[CD::S], const typename MP::T (&bm)[CD::M],
                                         uint32_t (&dw)[CD::NW]) {
    constexpr int B = CD::B, S = CD::S, PB = PH % B;
#pragma unroll
    for (int w = 0; w < CD::NW; ++w) dw[w] = 0;
#pragma unroll
    for (int a = 0; a < S; ++a) {
        if ((a >> PB) & 1) continue;                 // a holds the even state of a butterfly
        const int b = a | (1 << PB);                 // partner: the odd state
        const uint32_t u = rotr_bits(a, PH, B);      // logical 2m
        const uint32_t v = u + 1;                    // logical 2m+1
        const uint32_t j0 = u >> 1, j1 = j0 + S / 2; // successors (input 0 / input 1)
        const typename MP::T pe = r[a], po = r[b];
        MP::acs(r[a], pe, po, bm[CD::label(u, 0)], bm[CD::label(v, 0)], dw[j0 >> 5], 1u << (j0 & 31));
        MP::acs(r[b], pe, po, bm[CD::label(u, 1)], bm[CD::label(v, 1)], dw[j1 >> 5], 1u << (j1 & 31));
    }
}

// ---- packed-pair variant of the soft step (sm_100a add.rn.f32x2 = SASS FADD2) -----------------
// Registers are held as 64-bit pairs (physical 2q, 2q+1).  One FADD2 forms two candidate metrics
// in ONE issue slot; each lane of the packed add is an ordinary IEEE fp32 `pm + bm`, so results
// stay bit-identical to the scalar recipe.  At phase 0 a pair is one butterfly's (even, odd)
// members; at other phases it holds the same-parity members of two butterflies whose partners sit
// in pair q ^ (1 << (pb-1)).
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// dst = min(ce, co); decision (co < ce, strict) is ADDED as the exact power of two `w` to a float
// accumulator.  Pipe placement on sm_100a (measured, profiles/r01c): FSETP and FMNMX go to the ALU
// pipe, the predicated FADD to either FMA pipe — unlike a predicated integer add, which shares the
// FMA-heavy pipe with FADD2 and saturated it.  Up to 24 decisions per accumulator stay exact.
__device__ __forceinline__ float sel_lt(float ce, float co, float& acc, float w) {
    float dst;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.f32 p, %3, %2;\n\t"
        "min.f32 %0, %2, %3;\n\t"
        "@p add.f32 %1, %1, %4;\n\t}"
        : "=f"(dst), "+f"(acc) : "f"(ce), "f"(co), "f"(w));
    return dst;
}
// decision j of a step lives in accumulator j / 24 with weight 2^(j % 24)
constexpr int kAccBits = 24;
VDEC_HD constexpr float acc_weight(uint32_t j) { return (float)(1u << (j % kAccBits)); }

// float accumulators -> the survivor words (bit j of the step's S-bit decision vector = state j)
template <int S>
__device__ __forceinline__ void acc_to_words(const float (&da)[(S + kAccBits - 1) / kAccBits], uint32_t (&dw)[(S + 31) / 32]) {
    constexpr int NA = (S + kAccBits - 1) / kAccBits, NW = (S + 31) / 32;
    uint32_t u[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) u[a] = __float2uint_rz(da[a]);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        uint32_t word = 0;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            constexpr int dummy = 0; (void)dummy;
            const int sh = kAccBits * a - 32 * w;
            if (sh >= 0 && sh < 32) word |= u[a] << sh;
            else if (sh < 0 && sh > -kAccBits) word |= u[a] >> (-sh);
        }
        dw[w] = word;
    }
}

// Which (bm[x], bm[y]) operand pairs does phase PH of this code need?  Bit x*M+y of the result.
template <class CD, int PH>
VDEC_HD constexpr uint64_t used_pairs_mask() {
    constexpr int B = CD::B, S = CD::S, PB = PH % B, M = CD::M;
    uint64_t m = 0;
    for (int q = 0; q < S / 2; ++q) {
        if (PB == 0) {
            const uint32_t u = rotr_bits(2 * q, PH, B), v = u + 1;
            for (uint32_t i = 0; i < 2; ++i) m |= 1ull << (CD::label(u, i) * M + CD::label(v, i));
        } else {
            if ((q >> (PB - 1)) & 1) continue;
            const uint32_t u0 = rotr_bits(2 * q, PH, B), u1 = rotr_bits(2 * q + 1, PH, B);
            for (uint32_t i = 0; i < 2; ++i) {
                m |= 1ull << (CD::label(u0, i) * M + CD::label(u1, i));
                m |= 1ull << (CD::label(u0 + 1, i) * M + CD::label(u1 + 1, i));
            }
        }
    }
    return m;
}

// the same question for label alphabets larger than 8 (M*M > 64): is operand pair I = x*M + y needed at phase PH?
template <class CD, int PH>
VDEC_HD constexpr bool pair_used(int I) {
    constexpr int B = CD::B, S = CD::S, PB = PH % B, M = CD::M;
    for (int q = 0; q < S / 2; ++q) {
        if (PB == 0) {
            const uint32_t u = rotr_bits(2 * q, PH, B), v = u + 1;
            for (uint32_t i = 0; i < 2; ++i) if ((int)(CD::label(u, i) * M + CD::label(v, i)) == I) return true;
        } else {
            if ((q >> (PB - 1)) & 1) continue;
            const uint32_t u0 = rotr_bits(2 * q, PH, B), u1 = rotr_bits(2 * q + 1, PH, B);
            for (uint32_t i = 0; i < 2; ++i) {
                if ((int)(CD::label(u0, i) * M + CD::label(u1, i)) == I) return true;
                if ((int)(CD::label(u0 + 1, i) * M + CD::label(u1 + 1, i)) == I) return true;
            }
        }
    }
    return false;
}

// ---- decisions from the SIGN of packed differences (64-state steps) ---------------------------------------------------
// At every phase but 0 the two lanes of a candidate pair are the even (resp. odd) candidates of TWO new states, so one
// more FADD2 gives both differences co - ce, whose sign bits ARE the reference's strict '<' decisions: negative iff
// co < ce; equal candidates give +0 (tie -> even predecessor); INF - INF gives the positive canonical NaN (both
// unreachable -> 0, as `INF < INF` is false).  One PRMT copies the two sign bits, replicated over a byte each, into
// two byte lanes of a word and one LOP3 ORs the selected bit of each lane into the survivor word: 1.5 issue slots per
// decision pair-half instead of FSETP + predicated FADD (2.0), and no float-to-word conversion at the end of the step.
// Layout: the decision of new state j at rotation phase ph is bit dec_pos(j, ph) = rotl6(j, (ph + 4) % 6) of the
// step's 64-bit vector — the rotation puts the one bit in which the two states of a pair differ at position 3, i.e.
// into the byte-lane index, for every phase (phase 0 pairs j with j + 32).  The traceback applies the same rotation.
#ifndef VDEC_DEC_SIGN
#define VDEC_DEC_SIGN 1
#endif
// Measured (round-1 A/B builds; repeated in round 2, profiles/r02c_ab_operand_fetch_and_sign_decisions.txt): K = 7 (171,133) reference labels 107.3 -> 110.0 Gbit/s
// (1503 -> 1181 instructions per 6 steps), tile kernels K = 9 +3 %, K = 15 +2 %; the linear K = 7 code (-1.6 %) and
// rate 1/3 (-5 %) lose, so a code opts in through kSignDecisions.
template <class CD> constexpr bool kDecSign = VDEC_DEC_SIGN && CD::B == 6 && CD::S == 64 && CD::kSignDecisions;
VDEC_HD constexpr uint32_t dec_rot(int ph) { return (uint32_t)((ph % 6 + 4) % 6); }
VDEC_HD constexpr uint32_t dec_pos(uint32_t j, int ph) { return rotl_bits(j, dec_rot(ph), 6); }

__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
    uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// survivor bits PA and PB_ (same word, different byte lanes) <- sign bits of da and db
template <uint32_t PA, uint32_t PB_>
__device__ __forceinline__ void put_signs(uint32_t (&dw)[2], float da, float db) {
    static_assert((PA >> 5) == (PB_ >> 5) && ((PA >> 3) & 3u) != ((PB_ >> 3) & 3u), "one word, two byte lanes");
    constexpr uint32_t la = (PA >> 3) & 3u, lb = (PB_ >> 3) & 3u;
    constexpr uint32_t sel = (0xBu << (4 * la)) | (0xFu << (4 * lb));       // byte 3 of da / of db, sign-replicated
    constexpr uint32_t mask = (1u << (PA & 31u)) | (1u << (PB_ & 31u));
    uint32_t p;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(p) : "r"(__float_as_uint(da)), "r"(__float_as_uint(db)), "n"(sel));
    dw[PA >> 5] |= p & mask;
}
__device__ __forceinline__ float min_f32(float a, float b) {
    float r; asm("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}

template <class CD, int PH>
__device__ __forceinline__ void acs_step_f2(uint64_t (&rp)[CD::S / 2], const uint64_t (&bp)[CD::M * CD::M], uint32_t (&dw)[CD::NW]) {
    constexpr int B = CD::B, S = CD::S, PB = PH % B, M = CD::M, NA = (S + kAccBits - 1) / kAccBits;
    if constexpr (kDecSign<CD>) {
        dw[0] = 0; dw[1] = 0;
        if constexpr (PB == 0) {
            static_for<0, S / 2>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                constexpr uint32_t u = rotr_bits(2 * q, PH, B), v = u + 1;
                constexpr uint32_t j0 = u >> 1, j1 = j0 + S / 2;
                float e0, o0, e1, o1;
                unpack2(fadd2(rp[q], bp[CD::label(u, 0) * M + CD::label(v, 0)]), e0, o0);
                unpack2(fadd2(rp[q], bp[CD::label(u, 1) * M + CD::label(v, 1)]), e1, o1);
                put_signs<dec_pos(j0, PH), dec_pos(j1, PH)>(dw, __fsub_rn(o0, e0), __fsub_rn(o1, e1));
                rp[q] = pack2(min_f32(e0, o0), min_f32(e1, o1));
            });
        } else {
            static_for<0, S / 2>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                if constexpr (((q >> (PB - 1)) & 1) == 0) {
                    constexpr int q2 = q | (1 << (PB - 1));
                    constexpr uint32_t u0 = rotr_bits(2 * q, PH, B), u1 = rotr_bits(2 * q + 1, PH, B);
                    constexpr uint32_t v0 = u0 + 1, v1 = u1 + 1;
                    constexpr uint32_t ja = u0 >> 1, jb = u1 >> 1;          // input-0 successors; +S/2 for input 1
                    const uint64_t e = fadd2(rp[q],  bp[CD::label(u0, 0) * M + CD::label(u1, 0)]);
                    const uint64_t o = fadd2(rp[q2], bp[CD::label(v0, 0) * M + CD::label(v1, 0)]);
                    const uint64_t f = fadd2(rp[q],  bp[CD::label(u0, 1) * M + CD::label(u1, 1)]);
                    const uint64_t p = fadd2(rp[q2], bp[CD::label(v0, 1) * M + CD::label(v1, 1)]);
                    float ea, eb, oa, ob, fa, fb, pa, pb_, da0, db0, da1, db1;
                    unpack2(e, ea, eb); unpack2(o, oa, ob); unpack2(f, fa, fb); unpack2(p, pa, pb_);
                    unpack2(fsub2(o, e), da0, db0);
                    unpack2(fsub2(p, f), da1, db1);
                    put_signs<dec_pos(ja, PH), dec_pos(jb, PH)>(dw, da0, db0);
                    put_signs<dec_pos(ja + S / 2, PH), dec_pos(jb + S / 2, PH)>(dw, da1, db1);
                    rp[q] = pack2(min_f32(ea, oa), min_f32(eb, ob));
                    rp[q2] = pack2(min_f32(fa, pa), min_f32(fb, pb_));
                }
            });
        }
        return;
    }
    float da[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) da[a] = 0.f;
    if constexpr (PB == 0) {
        static_for<0, S / 2>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr uint32_t u = rotr_bits(2 * q, PH, B), v = u + 1;
            constexpr uint32_t j0 = u >> 1, j1 = j0 + S / 2;
            float e0, o0, e1, o1;
            unpack2(fadd2(rp[q], bp[CD::label(u, 0) * M + CD::label(v, 0)]), e0, o0);
            unpack2(fadd2(rp[q], bp[CD::label(u, 1) * M + CD::label(v, 1)]), e1, o1);
            const float n0 = sel_lt(e0, o0, da[j0 / kAccBits], acc_weight(j0));
            const float n1 = sel_lt(e1, o1, da[j1 / kAccBits], acc_weight(j1));
            rp[q] = pack2(n0, n1);
        });
    } else {
        static_for<0, S / 2>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            if constexpr (((q >> (PB - 1)) & 1) == 0) {
                constexpr int q2 = q | (1 << (PB - 1));
                constexpr uint32_t u0 = rotr_bits(2 * q, PH, B), u1 = rotr_bits(2 * q + 1, PH, B);
                constexpr uint32_t v0 = u0 + 1, v1 = u1 + 1;
                constexpr uint32_t ja = u0 >> 1, jb = u1 >> 1;          // input-0 successors; +S/2 for input 1
                float ea, eb, oa, ob, fa, fb, pa, pb_;
                unpack2(fadd2(rp[q],  bp[CD::label(u0, 0) * M + CD::label(u1, 0)]), ea, eb);
                unpack2(fadd2(rp[q2], bp[CD::label(v0, 0) * M + CD::label(v1, 0)]), oa, ob);
                unpack2(fadd2(rp[q],  bp[CD::label(u0, 1) * M + CD::label(u1, 1)]), fa, fb);
                unpack2(fadd2(rp[q2], bp[CD::label(v0, 1) * M + CD::label(v1, 1)]), pa, pb_);
                const float na = sel_lt(ea, oa, da[ja / kAccBits], acc_weight(ja));
                const float nb = sel_lt(eb, ob, da[jb / kAccBits], acc_weight(jb));
                const float ma = sel_lt(fa, pa, da[(ja + S / 2) / kAccBits], acc_weight(ja + S / 2));
                const float mb = sel_lt(fb, pb_, da[(jb + S / 2) / kAccBits], acc_weight(jb + S / 2));
                rp[q] = pack2(na, nb);
                rp[q2] = pack2(ma, mb);
            }
        });
    }
    acc_to_words<S>(da, dw);
}

template <int NW> struct SurvWord;
template <> struct SurvWord<1> { using type = uint32_t; };
template <> struct SurvWord<2> { using type = uint2; };

// survivor words are written once and read back ~0.7 ms later from L2 / HBM: VDEC_SURV_ST picks the store's cache
// operator (0 = default write-back, 1 = .cg: L2 only, 2 = .cs: streaming) so that they do not displace the input
// sectors that the ring's cp.async.ca keeps in L1
#ifndef VDEC_SURV_ST
#define VDEC_SURV_ST 1
#endif
template <int NW>
__device__ __forceinline__ void surv_store(uint32_t* p, const uint32_t (&dw)[NW]) {
    if constexpr (NW == 1) {
        if constexpr (VDEC_SURV_ST == 1) __stcg(p, dw[0]);
        else if constexpr (VDEC_SURV_ST == 2) __stcs(p, dw[0]);
        else *p = dw[0];
    } else {
        if constexpr (VDEC_SURV_ST == 1) __stcg(reinterpret_cast<uint2*>(p), make_uint2(dw[0], dw[1]));
        else if constexpr (VDEC_SURV_ST == 2) __stcs(reinterpret_cast<uint2*>(p), make_uint2(dw[0], dw[1]));
        else *reinterpret_cast<uint2*>(p) = make_uint2(dw[0], dw[1]);
    }
}

// lowest-index argmin over LOGICAL states when the registers are in rotation phase PH
template <class CD, class MP, int PH>
__device__ __forceinline__ void argmin_phase(const typename MP::T (&r)[CD::S], typename MP::T& best, uint32_t& idx) {
    best = MP::init_rest(); idx = 0;                                                           // :72-73
#pragma unroll
    for (int s = 0; s < CD::S; ++s) {
        const typename MP::T v = r[rotl_bits(s, PH, CD::B)];
        if (MP::less(v, best)) { best = v; idx = s; }                                          // :75-78
    }
}

// ---- input feeders -----------------------------------------------------------------------------
// Soft metrics arrive as [frame][step][M] floats: the 32 frames of a warp are 16 KB apart, so a
// per-step load touches 32 different cache lines.  Each lane therefore streams ITS OWN frame row into
// a private shared-memory ring with cp.async (LDGSTS, 16 B pieces, 128 B = one group at a time,
// two groups ahead of the consumer), and the ACS loop reads its step with one conflict-free LDS.128.
// No cross-lane sharing, hence no barrier: cp.async.wait_group is all the synchronisation needed.
// PB = bytes per cp.async piece.  PB = 16: one LDGSTS per 4 metrics, a lane's piece is 16 contiguous bytes, so
// the scalar operand loads of a warp hit 8 banks (4-way conflict).  PB = 4: one LDGSTS per metric (cp.async.ca),
// words interleaved across lanes, every operand load is conflict-free.  Measured (round-1 A/B build): the four
// times more LDGSTS requests cost far more than the conflicts (n = 2: 107.5 -> 85.7 Gbit/s, n = 3: 68.0 -> 35.7),
// so PB = 16 is the default for both; PB = 4 stays selectable for A/B builds (VDEC_RING_PB_M4 / _M8).
template <int M, int B, int PB>
struct SoftRing {
    static constexpr int kStepBytes = M * 4;
    static constexpr int kPieces = kStepBytes / PB;            // pieces per step
    static constexpr int G = B;                                // steps per prefetch group = one rotation period,
                                                               // so the prefetch sits at phase 0 only (no per-step test)
    static constexpr int NG = (M == 8) ? VDEC_RING_NG_M8 : VDEC_RING_NG_M4;  // groups in the ring
    static constexpr int kSlots = G * NG;
    static constexpr size_t kWarpBytes = (size_t)kSlots * 32 * kStepBytes;   // K=7, n=2: 9 KB
    static_assert(M == 4 || M == 8, "soft ring feeder supports n = 2, 3");
    static_assert(PB == 16 || PB == 4, "piece size");

    static constexpr uint32_t kPieceStride = 32 * PB;                      // ring bytes of one piece (all 32 lanes)
    static constexpr uint32_t kGroupBytes = (uint32_t)G * kPieces * kPieceStride;   // ring bytes of one group

    uint32_t base;        // shared-window address of this lane's first piece
    uint32_t T;
    float one, negzero;   // 1.0f and -0.0f as run-time values (get_pairs)
    const float* nextp;   // first step of the next group to prefetch (this lane's frame row)
    uint32_t next_t;      // its step index
    uint32_t woff;        // ring byte offset where that group goes
    uint32_t roff;        // ring byte offset of the group being consumed

    template <int OFF_S, int OFF_G>
    static __device__ __forceinline__ void copy_piece(uint32_t dst, const float* src) {
        if constexpr (PB == 16 && (M == 8 ? VDEC_RING_CA_M8 : VDEC_RING_CA))
            asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 16;" ::"r"(dst), "l"(src), "n"(OFF_S), "n"(OFF_G) : "memory");
        else if constexpr (PB == 16)
            asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%3], 16;" ::"r"(dst), "l"(src), "n"(OFF_S), "n"(OFF_G) : "memory");
        else
            asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 4;" ::"r"(dst), "l"(src), "n"(OFF_S), "n"(OFF_G) : "memory");
    }
    // Everything is kept incrementally (pointer, step, ring offsets) and the copies of a full group use
    // immediate offsets, so a period pays G*kPieces LDGSTS plus a handful of adds for its input.
    __device__ __forceinline__ void prefetch_next() {
        const uint32_t dst = base + woff;
        if (next_t + G <= T) {
            static_for<0, G * kPieces>([&](auto ic) {
                constexpr int I = decltype(ic)::value;
                copy_piece<I * (int)kPieceStride, I * PB>(dst, nextp);
            });
        } else {
            static_for<0, G * kPieces>([&](auto ic) {
                constexpr int I = decltype(ic)::value, s = I / kPieces;
                if (next_t + s < T) copy_piece<I * (int)kPieceStride, I * PB>(dst, nextp);
            });
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        nextp += G * M; next_t += G;
        woff = (woff == (NG - 1) * kGroupBytes) ? 0u : woff + kGroupBytes;
    }
    __device__ __forceinline__ void init(void* warp_smem, uint32_t lane, const float* row, uint32_t steps) {
        base = (uint32_t)__cvta_generic_to_shared(warp_smem) + lane * PB;
        T = steps; nextp = row; next_t = 0; woff = 0; roff = 0;
        one = __uint_as_float(0x3f800000u | (steps >> 31));          // steps < 2^31: exactly 1.0f and -0.0f
        negzero = __uint_as_float(0x80000000u | (steps >> 31));
        static_for<0, NG - 1>([&](auto) { prefetch_next(); });
    }
    // start of a period: keep the ring two groups ahead, make the current group visible
    __device__ __forceinline__ void begin_period() {
        prefetch_next();
        asm volatile("cp.async.wait_group %0;" ::"n"(NG - 1) : "memory");
    }
    __device__ __forceinline__ void end_period() { roff = (roff == (NG - 1) * kGroupBytes) ? 0u : roff + kGroupBytes; }
    // ring byte offset (immediate) of metric x of step PH of the current group
    template <int PH>
    VDEC_HD static constexpr int elem_off(int x) { return ((PH * M + x) * 4 / PB) * (int)kPieceStride + ((PH * M + x) * 4) % PB; }
    // packed mode: fetch every operand pair (bm[x], bm[y]) that MASK names straight into a 64-bit
    // register pair with two scalar loads.  Volatile so that ptxas neither fuses them into a v4 load
    // (which pins four consecutive registers and costs two MOVs per FADD2 operand) nor drops them.
    // VDEC_OPERAND_V4 (default): the step's M metrics come in with one LDS.128 per four (a lane's piece is 16 contiguous
    // bytes: conflict-free, one wavefront per quarter warp) and the operand pairs are formed by register moves on the
    // FMA pipes, which have slack.  The scalar form (two LDS.32 per pair, 4-way bank-conflicted by the 16-byte piece
    // layout: 565 M of 761 M shared wavefronts were replays, profiles/r01y) stays selectable for A/B builds.
    template <uint64_t MASK, int PH>
    __device__ __forceinline__ void get_pairs(uint64_t (&bp)[M * M]) const {
        const uint32_t a = base + roff;
        if constexpr (VDEC_OPERAND_V4 && PB == 16) {
            float bm[M];
            static_for<0, M / 4>([&](auto qc) {
                constexpr int Q = decltype(qc)::value;
                asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                             : "=f"(bm[4 * Q]), "=f"(bm[4 * Q + 1]), "=f"(bm[4 * Q + 2]), "=f"(bm[4 * Q + 3])
                             : "r"(a), "n"(elem_off<PH>(4 * Q)));
            });
            // Each half of an operand pair is a COPY of a metric made by an exact identity on the FMA pipes,
            // x * 1.0 + (-0.0) = x for every x (one, negzero: run-time values the compiler cannot fold), written
            // straight into the aligned register pair FADD2 wants.  Plain moves out of the LDS.128 quad are
            // copy-propagated by ptxas into two IMAD.MOV in front of almost every FADD2 (+ 70 % instructions).
            static_for<0, M * M>([&](auto ic) {
                constexpr int I = decltype(ic)::value, x = I / M, y = I % M;
                if constexpr ((MASK >> I) & 1ull) {
                    float lo, hi;
                    asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(lo) : "f"(bm[x]), "f"(one), "f"(negzero));
                    asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(hi) : "f"(bm[y]), "f"(one), "f"(negzero));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(bp[I]) : "f"(lo), "f"(hi));
                } else {
                    bp[I] = 0;
                }
            });
            return;
        }
        static_for<0, M * M>([&](auto ic) {
            constexpr int I = decltype(ic)::value, x = I / M, y = I % M;
            if constexpr ((MASK >> I) & 1ull) {
                float lo, hi;
                asm volatile("ld.volatile.shared.f32 %0, [%1+%2];" : "=f"(lo) : "r"(a), "n"(elem_off<PH>(x)));
                asm volatile("ld.volatile.shared.f32 %0, [%1+%2];" : "=f"(hi) : "r"(a), "n"(elem_off<PH>(y)));
                asm("mov.b64 %0, {%1, %2};" : "=l"(bp[I]) : "f"(lo), "f"(hi));
            } else {
                bp[I] = 0;
            }
        });
    }
    __device__ __forceinline__ void drain() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
};

// Raw I/Q rows: 8 B per step.  Same private ring as SoftRing with 8-byte pieces (cp.async.ca), one LDGSTS per step
// as before; the consumer reads its step with one conflict-free LDS.64 and computes the M metrics in registers:
// per point 2 subtractions + 2 squares per distinct coordinate, one add and the 3-operation exact division per
// metric (vdec_demap.cuh) — 24 FMA-pipe operations per step for n = 2, against ~200 issue slots of ACS.
template <int N, int B>
struct IqRing {
    static constexpr int M = 1 << N;
    static constexpr int G = B, NG = 3, kSlots = G * NG;
    static constexpr uint32_t kPieceStride = 32 * 8;
    static constexpr uint32_t kGroupBytes = (uint32_t)G * kPieceStride;
    static constexpr size_t kWarpBytes = (size_t)kSlots * 32 * 8;            // K = 7: 4.5 KB

    uint32_t base, T;
    const float* nextp;
    uint32_t next_t, woff, roff;

    template <int OFF_S, int OFF_G>
    static __device__ __forceinline__ void copy_piece(uint32_t dst, const float* src) {
        asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 8;" ::"r"(dst), "l"(src), "n"(OFF_S), "n"(OFF_G) : "memory");
    }
    __device__ __forceinline__ void prefetch_next() {
        const uint32_t dst = base + woff;
        if (next_t + G <= T) {
            static_for<0, G>([&](auto ic) {
                constexpr int I = decltype(ic)::value;
                copy_piece<I * (int)kPieceStride, I * 8>(dst, nextp);
            });
        } else {
            static_for<0, G>([&](auto ic) {
                constexpr int I = decltype(ic)::value;
                if (next_t + I < T) copy_piece<I * (int)kPieceStride, I * 8>(dst, nextp);
            });
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        nextp += G * 2; next_t += G;
        woff = (woff == (NG - 1) * kGroupBytes) ? 0u : woff + kGroupBytes;
    }
    __device__ __forceinline__ void init(void* warp_smem, uint32_t lane, const float* row, uint32_t steps) {
        base = (uint32_t)__cvta_generic_to_shared(warp_smem) + lane * 8;
        T = steps; nextp = row; next_t = 0; woff = 0; roff = 0;
        prefetch_next();
        prefetch_next();
    }
    __device__ __forceinline__ void begin_period() {
        prefetch_next();
        asm volatile("cp.async.wait_group 2;" ::: "memory");
    }
    __device__ __forceinline__ void end_period() { roff = (roff == (NG - 1) * kGroupBytes) ? 0u : roff + kGroupBytes; }
    template <uint64_t MASK, int PH>
    __device__ __forceinline__ void get_pairs(uint64_t (&bp)[M * M]) const {
        float x, y;
        asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(x), "=f"(y) : "r"(base + roff), "n"(PH * (int)kPieceStride));
        float bm[M];
        demap_point<N>(x, y, bm);
        static_for<0, M * M>([&](auto ic) {
            constexpr int I = decltype(ic)::value, a = I / M, b = I % M;
            if constexpr ((MASK >> I) & 1ull) asm("mov.b64 %0, {%1, %2};" : "=l"(bp[I]) : "f"(bm[a]), "f"(bm[b]));
            else bp[I] = 0;
        });
    }
    __device__ __forceinline__ void drain() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
};

// Hard symbols are one byte per step; a warp's 32 frames are contiguous (32*T bytes), so plain
// read-only loads stay in L1.
template <int M, int B>
struct HardDirect {
    static constexpr size_t kWarpBytes = 0;
    const uint8_t* src;
    __device__ __forceinline__ void init(void*, uint32_t, const uint8_t* row, uint32_t) { src = row; }
    __device__ __forceinline__ void begin_period() const {}
    __device__ __forceinline__ void end_period() {}
    __device__ __forceinline__ void get(uint32_t t, uint32_t (&bm)[M]) const {
        const uint32_t r = __ldg(src + t) & (M - 1);                                          // :235-237
#pragma unroll
        for (int e = 0; e < M; ++e) bm[e] = __popc(e ^ r);                                   // :69-72
    }
    __device__ __forceinline__ void drain() const {}
};

template <class MP, int M, int B> struct FeederOf;
#ifndef VDEC_RING_PB_M4
#define VDEC_RING_PB_M4 16
#endif
#ifndef VDEC_RING_PB_M8
#define VDEC_RING_PB_M8 16
#endif
template <int M, int B> struct FeederOf<SoftF32, M, B> { using type = SoftRing<M, B, (M == 8 ? VDEC_RING_PB_M8 : VDEC_RING_PB_M4)>; };
template <int M, int B> struct FeederOf<HardU16, M, B> { using type = HardDirect<M, B>; };
template <int M, int B> struct FeederOf<SoftIq, M, B> { using type = IqRing<(M == 2 ? 1 : M == 4 ? 2 : 3), B>; };

#ifndef VDEC_TRACE_CHUNK
#define VDEC_TRACE_CHUNK 16
#endif
constexpr int kTraceChunk = VDEC_TRACE_CHUNK;     // survivor words prefetched per traceback chunk
static_assert(kTraceChunk == 16, "traceback emits 16 output bits per 16-step chunk");

template <class CD, class MP>
__global__ void __launch_bounds__(kRegsThreads, (MP::kPacked ? CD::kMinBlocks : 4))
vit_regs_kernel(const typename MP::In* __restrict__ in, uint8_t* __restrict__ out, int32_t* __restrict__ path_metric,
                uint32_t* __restrict__ surv, uint32_t nframes, uint32_t T, uint32_t out_bytes, uint32_t stage_words,
                uint32_t stride_steps, uint32_t init_all_from) {
    using MT = typename MP::T;
    using Feeder = typename FeederOf<MP, CD::M, CD::B>::type;
    using SW = typename SurvWord<CD::NW>::type;
    constexpr int S = CD::S, B = CD::B, M = CD::M, NW = CD::NW;
    // dynamic shared memory: input rings [warps][kWarpBytes] during the forward pass, then reused as
    // the output staging tile [threads][stage_words] (each warp only ever touches its own part)
    extern __shared__ __align__(16) unsigned char s_raw[];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t f = blockIdx.x * kRegsThreads + tid;
    const bool live = f < nframes;
    const uint32_t fr = live ? f : nframes - 1;      // dead lanes shadow the last frame, store nothing
    const size_t step_elems = MP::template step_elems<M>();
    // survivor scratch: [global warp][step][lane][NW]
    uint32_t* sv = surv + ((size_t)(f >> 5) * T * 32 + lane) * NW;
    const size_t ring_bytes = Feeder::kWarpBytes, tile_bytes = (size_t)32 * stage_words * 4;
    const size_t warp_smem = ring_bytes > tile_bytes ? ring_bytes : tile_bytes;
    unsigned char* my_smem = s_raw + warp * warp_smem;

    // All CTAs of a wave start together, run the same number of steps and therefore reach their (latency-bound,
    // low-issue) traceback together: the SM then has nothing to overlap it with.  The CTAs that share an SM in the
    // FIRST wave are started a fraction of a CTA lifetime apart, once; the phase shift persists for the rest of the
    // grid, so that a traceback always runs next to three forward passes.
    if constexpr (CD::kStagger > 0 && MP::kPacked) {
        constexpr uint32_t kSlots = (uint32_t)CD::kMinBlocks;
        uint32_t nsm;
        asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        const uint32_t slot = blockIdx.x / nsm;                       // first wave: CTAs are dealt out one per SM, round after round
        if (slot > 0 && slot < kSlots && gridDim.x > 2 * kSlots * nsm) {
            const long long until = clock64() + (long long)slot * T * (CD::kStagger / (int)kSlots);
            while (clock64() < until) __nanosleep(2000);
        }
    }
    Feeder feed;
    feed.init(my_smem, lane, in + (size_t)fr * stride_steps * step_elems, T);

    MT r[S];
    r[0] = 0;                                        // decoder_reset: state 0 starts at metric 0 (:215)
    const MT rest = (f >= init_all_from) ? MT(0) : MP::init_rest();      // sliding-window blocks: start state unknown
#pragma unroll
    for (int s = 1; s < S; ++s) r[s] = rest;
    [[maybe_unused]] uint64_t rp[S / 2];             // packed mode: physical registers (2q, 2q+1) as one pair
    if constexpr (MP::kPacked) {
#pragma unroll
        for (int q = 0; q < S / 2; ++q) rp[q] = pack2(r[2 * q], r[2 * q + 1]);
    }

    uint32_t dw[NW];
    MT bm[M];

    // ---- forward pass: full periods of B steps, then the T % B remainder --------------------
    uint32_t t = 0;
    for (; t + B <= T; t += B) {
        feed.begin_period();
        static_for<0, B>([&](auto phc) {
            constexpr int PH = decltype(phc)::value;
            if constexpr (MP::kPacked) {
                uint64_t bp[M * M];
                feed.template get_pairs<used_pairs_mask<CD, PH>(), PH>(bp);
                acs_step_f2<CD, PH>(rp, bp, dw);
            } else {
                feed.get(t + PH, bm);
                acs_step<CD, MP, PH>(r, bm, dw);
            }
            if (live) surv_store<NW>(sv + (size_t)(t + PH) * 32 * NW, dw);
        });
        feed.end_period();
    }
    const uint32_t rem = T - t;                       // 0 .. B-1, warp-uniform
    if (rem) feed.begin_period();
    static_for<0, B - 1>([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (PH < (int)rem) {
            if constexpr (MP::kPacked) {
                uint64_t bp[M * M];
                feed.template get_pairs<used_pairs_mask<CD, PH>(), PH>(bp);
                acs_step_f2<CD, PH>(rp, bp, dw);
            } else {
                feed.get(t + PH, bm);
                acs_step<CD, MP, PH>(r, bm, dw);
            }
            if (live) surv_store<NW>(sv + (size_t)(t + PH) * 32 * NW, dw);
        }
    });
    feed.drain();
    if constexpr (MP::kPacked) {
#pragma unroll
        for (int q = 0; q < S / 2; ++q) unpack2(rp[q], r[2 * q], r[2 * q + 1]);
    }

    // ---- final state: lowest-index argmin over logical states (:72-79) ----------------------
    MT best = MP::init_rest(); uint32_t cur = 0;
    static_for<0, B>([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (PH == (int)rem) argmin_phase<CD, MP, PH>(r, best, cur);
    });
    if (live && path_metric != nullptr) path_metric[f] = MP::to_path_metric(best);            // :94, :143
    __syncwarp();                                     // ring reads are done before the tile is reused

    // ---- traceback (:81-89): decision of state `cur` at step tt, decoded bit = MSB of cur -----
    // The survivor words of a step do not depend on the state, so they are fetched a chunk ahead
    // (coalesced: 32 lanes x NW words).  The state is kept as an unmasked shift register `h`
    // (cur = h & (S-1)): the decoded bit of step tt is the decision taken B-1 steps later on the way
    // back, so bit B-1+i of h at step tt is the decoded bit of step tt+i and 16 output bits fall out
    // of h every 16 steps without any per-step accumulation.  Rows are staged at their final out_bytes
    // pitch, so the warp's tile is the byte image of its 32 contiguous output rows.
    unsigned char* tile = my_smem;
    if (live) {
        // Two register sets (A, B) of 16 steps each: the words of chunk c-1 are requested before chunk c is
        // walked, and an L2 prefetch runs three chunks further ahead (one 128-byte line per lane), so the walk
        // never waits for HBM.  For S = 64 the word of a step (x: states 0..31, y: 32..63) is chosen by bit 5 of
        // cur, i.e. by the decision taken 6 steps earlier on the way back: the select is done then (`sa`, `sb`)
        // and the dependent chain of a step is one variable shift and one 3-input logic op.
        const SW* svw = reinterpret_cast<const SW*>(sv);
        unsigned char* srow = tile + (size_t)lane * out_bytes;
        SW wa[kTraceChunk], wb[kTraceChunk];
        [[maybe_unused]] uint32_t sa[kTraceChunk], sb[kTraceChunk];
        int base = ((int)T - 1) & ~(kTraceChunk - 1);
        const int jmax = ((int)T - 1) & (kTraceChunk - 1);
        uint32_t h = cur;
        auto l2_prefetch = [&](int cb) {          // chunk starting at step cb: 16 * NW lines, one per lane
            if (cb >= 0 && lane < 16u * NW)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(svw) - (size_t)lane * NW * 4 +
                                                               (size_t)cb * 128 * NW + (size_t)lane * 128));
        };
        auto load = [&](SW (&w)[kTraceChunk], int cb) {
            if (cb >= 0) {
#pragma unroll
                for (int j = 0; j < kTraceChunk; ++j) w[j] = __ldcg(svw + (size_t)(cb + j) * 32);
            }
        };
#pragma unroll
        for (int j = 0; j < kTraceChunk; ++j)
            if (j <= jmax) wa[j] = __ldcg(svw + (size_t)(base + j) * 32);
        load(wb, base - kTraceChunk);
        l2_prefetch(base - 2 * kTraceChunk);
        l2_prefetch(base - 3 * kTraceChunk);
        if constexpr (NW == 2 && !(MP::kPacked && kDecSign<CD>)) {   // the last 6 steps take their word-select bit from the final state itself
#pragma unroll
            for (int j = 0; j < kTraceChunk; ++j) {
                const int ia = jmax - j, ib = jmax + kTraceChunk - j;       // steps back from T-1
                if (ia >= 0 && ia <= 5) sa[j] = ((cur >> (5 - ia)) & 1u) ? wa[j].y : wa[j].x;
                if (ib <= 5) sb[j] = ((cur >> (5 - ib)) & 1u) ? wb[j].y : wb[j].x;
            }
        }
        // Sign-based decisions (acs_step_f2): the bit of state cur at step tt is jp = rotl6(cur, (tt + 4) % 6).  One step
        // back the state becomes (cur << 1 | d) & 63 and the rotation one less, which leaves every bit of jp where it is
        // except bit (tt + 3) % 6, which takes the decision d: the index is kept incrementally with one bit insert per
        // step.  The bit position is a compile-time constant when the phase of a chunk is: chunks are 16 steps, phases
        // have period 6, so the walk is instantiated for the three residues of (chunk start / 16) mod 3.
        constexpr bool kSignLayout = MP::kPacked && kDecSign<CD>;
        [[maybe_unused]] uint32_t jp = (((cur & 63u) * 65u) >> (6u - (T + 3u) % 6u)) & 63u;      // for tt = T - 1
        auto walk = [&](SW (&cw)[kTraceChunk], uint32_t (&cs)[kTraceChunk], SW (&nw)[kTraceChunk], uint32_t (&ns)[kTraceChunk], auto guard,
                        auto cls) {
            constexpr bool GUARD = decltype(guard)::value;      // first chunk only: steps >= T do not exist
            [[maybe_unused]] constexpr int CLS = decltype(cls)::value;
#pragma unroll
            for (int j = kTraceChunk - 1; j >= 0; --j) {
                const int tt = base + j;
                if (GUARD && j > jmax) continue;
                if constexpr (kSignLayout) {
                    if ((j & 15) == 0) {
                        const uint32_t rv = __brev(h >> (B - 1));
                        srow[tt >> 3] = (unsigned char)(rv >> 24);
                        if (!GUARD || tt + 8 < (int)T) srow[(tt >> 3) + 1] = (unsigned char)(rv >> 16);
                    }
                    const uint32_t bpos = (uint32_t)((16 * CLS + j) % 6 + 3) % 6u;      // (tt + 3) % 6, a constant after unrolling
                    const uint32_t word = (jp & 32u) ? cw[j].y : cw[j].x;
                    const uint32_t x = __funnelshift_r(word, 0u, jp);                   // bit 0: decision of state cur at step tt
                    h = (h << 1) | (x & 1u);                                            // predecessor (:86)
                    jp = (jp & ~(1u << bpos)) | ((x << bpos) & (1u << bpos));
                    continue;
                }
                if ((j & 15) == 0) {                             // bits tt .. tt+15 -> bytes tt/8, tt/8+1, MSB-first (:88)
                    const uint32_t rv = __brev(h >> (B - 1));
                    srow[tt >> 3] = (unsigned char)(rv >> 24);
                    if (!GUARD || tt + 8 < (int)T) srow[(tt >> 3) + 1] = (unsigned char)(rv >> 16);
                }
                uint32_t word;
                if constexpr (NW == 1) word = cw[j];
                else word = cs[j];
                const uint32_t d = (word >> (h & (S - 1) & 31u)) & 1u;
                h = (h << 1) | d;                                               // predecessor (:86)
                if constexpr (NW == 2) {                         // d is bit 5 of cur six steps further back
                    if (j >= 6) cs[j - 6] = d ? cw[j - 6].y : cw[j - 6].x;
                    else ns[j + kTraceChunk - 6] = d ? nw[j + kTraceChunk - 6].y : nw[j + kTraceChunk - 6].x;
                }
            }
        };
        // chunk class = (base / 16) mod 3, warp-uniform; only the sign layout distinguishes the classes
        auto walk_cls = [&](SW (&cw)[kTraceChunk], uint32_t (&cs)[kTraceChunk], SW (&nw)[kTraceChunk], uint32_t (&ns)[kTraceChunk], auto guard) {
            if constexpr (!kSignLayout) {
                walk(cw, cs, nw, ns, guard, std::integral_constant<int, 0>{});
            } else {
                const uint32_t cls = ((uint32_t)base >> 4) % 3u;
                if (cls == 0) walk(cw, cs, nw, ns, guard, std::integral_constant<int, 0>{});
                else if (cls == 1) walk(cw, cs, nw, ns, guard, std::integral_constant<int, 1>{});
                else walk(cw, cs, nw, ns, guard, std::integral_constant<int, 2>{});
            }
        };
        walk_cls(wa, sa, wb, sb, std::true_type{});
        for (;;) {
            base -= kTraceChunk;
            if (base < 0) break;
            load(wa, base - kTraceChunk);
            l2_prefetch(base - 4 * kTraceChunk);
            walk_cls(wb, sb, wa, sa, std::false_type{});
            base -= kTraceChunk;
            if (base < 0) break;
            load(wb, base - kTraceChunk);
            l2_prefetch(base - 4 * kTraceChunk);
            walk_cls(wa, sa, wb, sb, std::false_type{});
        }
    }
    __syncwarp();

    // ---- write the warp's 32 output rows (contiguous in global memory) ----------------------
    const uint32_t wbase = (blockIdx.x * kRegsThreads + (tid & ~31u));     // first frame of this warp
    if (wbase < nframes) {
        const uint32_t rows = min(32u, nframes - wbase);
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
// ---- host side ------------------------------------------------------------------------------------
// Launch shape of vit_regs_kernel as a function of (K, n, metric) only, so that kernels compiled at run time
// (vdec_jit.cu) are launched by the same code as the built-in instantiations.
struct RegsShape {
    uint32_t nw;            // survivor words per step and frame
    size_t ring_bytes;      // input ring per warp
};
inline RegsShape regs_shape(uint32_t K, uint32_t n, uint32_t metric, bool iq = false) {
    const uint32_t S = 1u << (K - 1), M = 1u << n;
    const uint32_t ng = iq ? 3 : (M == 8 ? VDEC_RING_NG_M8 : VDEC_RING_NG_M4);      // SoftRing / IqRing: NG groups of K-1 steps
    return RegsShape{(S + 31) / 32, metric == VDEC_SOFT_F32 ? (size_t)(K - 1) * ng * 32 * (iq ? 8 : M * 4) : 0};
}

// `kern` is a __global__ symbol of this library or a cudaKernel_t of a run-time compiled module
inline int regs_launch_kernel(const void* kern, const RegsShape& sh, const LaunchGeom& g, const void* in, uint8_t* out,
                              int32_t* path_metric, uint32_t* surv, cudaStream_t st) {
    uint32_t stage_words = (g.out_bytes + 3) / 4;
    if ((stage_words & 1) == 0) ++stage_words;                 // odd stride: conflict-free column writes
    size_t warp_smem = (size_t)32 * stage_words * 4;
    if (sh.ring_bytes > warp_smem) warp_smem = sh.ring_bytes;
    const size_t smem = (size_t)(kRegsThreads / 32) * warp_smem;
    if (smem > 48 * 1024) VDEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (g.nframes + kRegsThreads - 1) / kRegsThreads;
    uint32_t nframes = g.nframes, T = g.T, out_bytes = g.out_bytes, stride = g.stride_steps, init_all_from = g.init_all_from;
    void* args[] = {&in, &out, &path_metric, &surv, &nframes, &T, &out_bytes, &stage_words, &stride, &init_all_from};
    VDEC_CUDA(cudaLaunchKernel(kern, dim3(blocks), dim3(kRegsThreads), args, smem, st));
    return 0;
}

inline size_t regs_surv_bytes_nw(const LaunchGeom& g, uint32_t nw) {
    return (size_t)((g.nframes + 31) / 32) * g.T * 32 * nw * sizeof(uint32_t);
}

template <class CD, class MP>
int regs_launch(const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* surv, cudaStream_t st) {
    using Feeder = typename FeederOf<MP, CD::M, CD::B>::type;
    constexpr RegsShape sh{(uint32_t)CD::NW, Feeder::kWarpBytes};
    static_assert(Feeder::kWarpBytes == (MP::kMetric == VDEC_SOFT_F32
                                             ? (size_t)CD::B * (kIsIq<MP> ? 3 : (CD::M == 8 ? VDEC_RING_NG_M8 : VDEC_RING_NG_M4)) * 32 * (kIsIq<MP> ? 8 : CD::M * 4) : 0),
                  "regs_shape() must describe the feeder");
    return regs_launch_kernel(reinterpret_cast<const void*>(&vit_regs_kernel<CD, MP>), sh, g, in, out, path_metric, surv, st);
}

template <class CD>
size_t regs_surv_bytes(const LaunchGeom& g) { return regs_surv_bytes_nw(g, CD::NW); }

template <class CD, class MP>
RegsEntry make_regs_entry(const char* name) {
    RegsEntry e{};
    e.name = name; e.K = CD::K; e.n = CD::N; e.metric = MP::kMetric;
    for (int s = 0; s < CD::S; ++s) { e.labels[2 * s] = (uint8_t)CD::label(s, 0); e.labels[2 * s + 1] = (uint8_t)CD::label(s, 1); }
    e.launch = &regs_launch<CD, MP>;
    if constexpr (std::is_same_v<MP, SoftF32>) e.launch_iq = &regs_launch<CD, SoftIq>;      // raw-I/Q entry, metrics formed in the feeder
    e.surv_bytes = &regs_surv_bytes<CD>;
    return e;
}
#endif  // !__CUDACC_RTC__

}  // namespace vd
