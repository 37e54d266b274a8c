// vdec_regs.cu — registry of code-specialised register kernels (see vdec_regs.cuh).
//
// Each entry is compiled for one trellis labelling.  A request is served by an entry when the
// label table computed from the caller's polynomials (host side, vdec_api.cu) equals the entry's
// table, so e.g. codes without a D^3 tap share one entry for both parity modes.
#include <string.h>

#include "vdec_regs.cuh"

namespace vd {

// K=7 (171,133)_8 rate 1/2 — BASELINE config 2 (headline)
using K7R12Ref = CodeDef<7, 2, 0xF200000000000000ull, 0xB600000000000000ull, 0, true>;
using K7R12Lin = CodeDef<7, 2, 0xF200000000000000ull, 0xB600000000000000ull, 0, false>;
// K=7 (133,171,165)_8 rate 1/3 — BASELINE config 3
using K7R13Ref = CodeDef<7, 3, 0xB600000000000000ull, 0xF200000000000000ull, 0xEA00000000000000ull, true>;
using K7R13Lin = CodeDef<7, 3, 0xB600000000000000ull, 0xF200000000000000ull, 0xEA00000000000000ull, false>;
// common/codebook.c entries 0,1,2,3,5 (index 4 is the K=32 WSPR code: not Viterbi-decodable)
using CB0 = CodeDef<3, 2, 0xA000000000000000ull, 0x6000000000000000ull, 0, true>;   // codebook.c:14-17
using CB1 = CodeDef<4, 2, 0xB000000000000000ull, 0xE000000000000000ull, 0, true>;   // :21-24
using CB2 = CodeDef<5, 2, 0xA800000000000000ull, 0xF000000000000000ull, 0, true>;   // :28-31
using CB3 = CodeDef<6, 2, 0xB400000000000000ull, 0xE800000000000000ull, 0, true>;   // :35-38
using CB5 = CodeDef<3, 3, 0xA000000000000000ull, 0xC000000000000000ull, 0x2000000000000000ull, true>;   // :50-54
// textbook (true parity) labels of the codebook entries that have a D^3 tap (0 and 5 have none: one table serves both modes)
using CB1L = CodeDef<4, 2, 0xB000000000000000ull, 0xE000000000000000ull, 0, false>;
using CB2L = CodeDef<5, 2, 0xA800000000000000ull, 0xF000000000000000ull, 0, false>;
using CB3L = CodeDef<6, 2, 0xB400000000000000ull, 0xE800000000000000ull, 0, false>;

static const RegsEntry g_entries[] = {
    make_regs_entry<K7R12Ref, SoftF32>("regs_soft_f32<K7 (171,133) ref-labels>"),
    make_regs_entry<K7R12Lin, SoftF32>("regs_soft_f32<K7 (171,133) linear>"),
    make_regs_entry<K7R13Ref, SoftF32>("regs_soft_f32<K7 (133,171,165) ref-labels>"),
    make_regs_entry<K7R13Lin, SoftF32>("regs_soft_f32<K7 (133,171,165) linear>"),
    make_regs_entry<CB0, SoftF32>("regs_soft_f32<codebook0 K3>"),
    make_regs_entry<CB1, SoftF32>("regs_soft_f32<codebook1 K4>"),
    make_regs_entry<CB2, SoftF32>("regs_soft_f32<codebook2 K5>"),
    make_regs_entry<CB3, SoftF32>("regs_soft_f32<codebook3 K6>"),
    make_regs_entry<CB5, SoftF32>("regs_soft_f32<codebook5 K3 r1/3>"),
    make_regs_entry<CB1L, SoftF32>("regs_soft_f32<codebook1 K4 linear>"),
    make_regs_entry<CB2L, SoftF32>("regs_soft_f32<codebook2 K5 linear>"),
    make_regs_entry<CB3L, SoftF32>("regs_soft_f32<codebook3 K6 linear>"),
    make_regs_entry<K7R12Ref, HardU16>("regs_hard_u16<K7 (171,133) ref-labels>"),
    make_regs_entry<K7R12Lin, HardU16>("regs_hard_u16<K7 (171,133) linear>"),
    make_regs_entry<CB0, HardU16>("regs_hard_u16<codebook0 K3>"),
    make_regs_entry<CB1, HardU16>("regs_hard_u16<codebook1 K4>"),
    make_regs_entry<CB2, HardU16>("regs_hard_u16<codebook2 K5>"),
    make_regs_entry<CB3, HardU16>("regs_hard_u16<codebook3 K6>"),
    make_regs_entry<CB5, HardU16>("regs_hard_u16<codebook5 K3 r1/3>"),
    make_regs_entry<CB1L, HardU16>("regs_hard_u16<codebook1 K4 linear>"),
    make_regs_entry<CB2L, HardU16>("regs_hard_u16<codebook2 K5 linear>"),
    make_regs_entry<CB3L, HardU16>("regs_hard_u16<codebook3 K6 linear>"),
};

int regs_count() { return (int)(sizeof(g_entries) / sizeof(g_entries[0])); }
const RegsEntry* regs_at(int i) { return &g_entries[i]; }

const RegsEntry* regs_find(uint32_t K, uint32_t n, uint32_t metric, const uint8_t* labels) {
    if (K > 7) return nullptr;
    const uint32_t S = 1u << (K - 1);
    for (const RegsEntry& e : g_entries)
        if (e.K == K && e.n == n && e.metric == metric && memcmp(e.labels, labels, 2 * S) == 0) return &e;
    return nullptr;
}

}  // namespace vd
