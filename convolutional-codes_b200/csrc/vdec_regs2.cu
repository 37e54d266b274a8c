// vdec_regs2.cu — registry of packed 16-bit (two frames per thread) kernels, see vdec_regs2.cuh.
#include <string.h>

#include "vdec_regs2.cuh"

namespace vd {

using K7R12Ref = CodeDef<7, 2, 0xF200000000000000ull, 0xB600000000000000ull, 0, true>;
using K7R12Lin = CodeDef<7, 2, 0xF200000000000000ull, 0xB600000000000000ull, 0, false>;
using K7R13Ref = CodeDef<7, 3, 0xB600000000000000ull, 0xF200000000000000ull, 0xEA00000000000000ull, true>;
using K7R13Lin = CodeDef<7, 3, 0xB600000000000000ull, 0xF200000000000000ull, 0xEA00000000000000ull, false>;
using CB0 = CodeDef<3, 2, 0xA000000000000000ull, 0x6000000000000000ull, 0, true>;   // common/codebook.c:14-17
using CB1 = CodeDef<4, 2, 0xB000000000000000ull, 0xE000000000000000ull, 0, true>;   // :21-24
using CB2 = CodeDef<5, 2, 0xA800000000000000ull, 0xF000000000000000ull, 0, true>;   // :28-31
using CB3 = CodeDef<6, 2, 0xB400000000000000ull, 0xE800000000000000ull, 0, true>;   // :35-38
using CB5 = CodeDef<3, 3, 0xA000000000000000ull, 0xC000000000000000ull, 0x2000000000000000ull, true>;   // :50-54
using CB1L = CodeDef<4, 2, 0xB000000000000000ull, 0xE000000000000000ull, 0, false>;  // true-parity labels (D^3 tap present)
using CB2L = CodeDef<5, 2, 0xA800000000000000ull, 0xF000000000000000ull, 0, false>;
using CB3L = CodeDef<6, 2, 0xB400000000000000ull, 0xE800000000000000ull, 0, false>;

static const Regs2Entry g_entries2[] = {
    make_regs2_entry<K7R12Ref, kModeHard>("regs2_hard_u16x2<K7 (171,133) ref-labels>"),
    make_regs2_entry<K7R12Lin, kModeHard>("regs2_hard_u16x2<K7 (171,133) linear>"),
    make_regs2_entry<K7R13Ref, kModeHard>("regs2_hard_u16x2<K7 (133,171,165) ref-labels>"),
    make_regs2_entry<K7R13Lin, kModeHard>("regs2_hard_u16x2<K7 (133,171,165) linear>"),
    make_regs2_entry<CB0, kModeHard>("regs2_hard_u16x2<codebook0 K3>"),
    make_regs2_entry<CB1, kModeHard>("regs2_hard_u16x2<codebook1 K4>"),
    make_regs2_entry<CB2, kModeHard>("regs2_hard_u16x2<codebook2 K5>"),
    make_regs2_entry<CB3, kModeHard>("regs2_hard_u16x2<codebook3 K6>"),
    make_regs2_entry<CB5, kModeHard>("regs2_hard_u16x2<codebook5 K3 r1/3>"),
    make_regs2_entry<CB1L, kModeHard>("regs2_hard_u16x2<codebook1 K4 linear>"),
    make_regs2_entry<CB2L, kModeHard>("regs2_hard_u16x2<codebook2 K5 linear>"),
    make_regs2_entry<CB3L, kModeHard>("regs2_hard_u16x2<codebook3 K6 linear>"),
    make_regs2_entry<K7R12Ref, kModeQuant>("regs2_soft_u8x2<K7 (171,133) ref-labels>"),
    make_regs2_entry<K7R12Lin, kModeQuant>("regs2_soft_u8x2<K7 (171,133) linear>"),
    make_regs2_entry<CB0, kModeQuant>("regs2_soft_u8x2<codebook0 K3>"),
    make_regs2_entry<CB1, kModeQuant>("regs2_soft_u8x2<codebook1 K4>"),
    make_regs2_entry<CB2, kModeQuant>("regs2_soft_u8x2<codebook2 K5>"),
    make_regs2_entry<CB3, kModeQuant>("regs2_soft_u8x2<codebook3 K6>"),
    make_regs2_entry<CB1L, kModeQuant>("regs2_soft_u8x2<codebook1 K4 linear>"),
    make_regs2_entry<CB2L, kModeQuant>("regs2_soft_u8x2<codebook2 K5 linear>"),
    make_regs2_entry<CB3L, kModeQuant>("regs2_soft_u8x2<codebook3 K6 linear>"),
};

const Regs2Entry* regs2_find(uint32_t K, uint32_t n, uint32_t mode, const uint8_t* labels) {
    if (K > 7) return nullptr;
    const uint32_t S = 1u << (K - 1);
    for (const Regs2Entry& e : g_entries2)
        if (e.K == K && e.n == n && e.mode == mode && memcmp(e.labels, labels, 2 * S) == 0) return &e;
    return nullptr;
}

const char* regs2_name(const Regs2Entry* e) { return e->name; }
int regs2_run(const Regs2Entry* e, const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric,
              uint32_t* scratch, cudaStream_t st) { return e->launch(g, in, out, path_metric, scratch, st); }
size_t regs2_bytes(const Regs2Entry* e, const LaunchGeom& g) { return e->scratch_bytes(g); }

}  // namespace vd
