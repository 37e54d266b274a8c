// vdec_tile2.cu — registry of packed 16-bit register-tile kernels for K = 8 .. 15 (see vdec_tile2.cuh):
// hard decisions and quantised soft metrics on large trellises.
#include <string.h>

#include "vdec_tile2.cuh"

namespace vd {

using T8   = TileDef<8,  0xF900000000000000ull, 0xA700000000000000ull>;          // (371,247)_8
using T8R  = TileDef<8,  0xF900000000000000ull, 0xA700000000000000ull, true>;
using T9   = TileDef<9,  0xB880000000000000ull, 0xF580000000000000ull>;          // (561,753)_8
using T9R  = TileDef<9,  0xB880000000000000ull, 0xF580000000000000ull, true>;
using T11  = TileDef<11, 0x9BA0000000000000ull, 0xF620000000000000ull>;          // (2335,3661)_8
using T13  = TileDef<13, 0x8AD8000000000000ull, 0xFD88000000000000ull>;          // (10533,17661)_8
using T15  = TileDef<15, 0x99A2000000000000ull, 0xA572000000000000ull>;          // (46321,51271)_8

static const Tile2Entry* entries(int* count) {
    static const Tile2Entry e[] = {
        make_tile2_entry<T8, kModeHard>("tile2_hard_u16x2<K8 (371,247) linear, 2 threads/pair>"),
        make_tile2_entry<T8R, kModeHard>("tile2_hard_u16x2<K8 (371,247) ref-labels, 2 threads/pair>"),
        make_tile2_entry<T9, kModeHard>("tile2_hard_u16x2<K9 (561,753) linear, 4 threads/pair>"),
        make_tile2_entry<T9R, kModeHard>("tile2_hard_u16x2<K9 (561,753) ref-labels, 4 threads/pair>"),
        make_tile2_entry<T11, kModeHard>("tile2_hard_u16x2<K11 (2335,3661) linear, 16 threads/pair>"),
        make_tile2_entry<T13, kModeHard>("tile2_hard_u16x2<K13 (10533,17661) linear, 64 threads/pair>"),
        make_tile2_entry<T15, kModeHard>("tile2_hard_u16x2<K15 (46321,51271) linear, 256 threads/pair>"),
        make_tile2_entry<T8R, kModeQuant>("tile2_soft_u8x2<K8 (371,247) ref-labels, 2 threads/pair>"),
        make_tile2_entry<T9, kModeQuant>("tile2_soft_u8x2<K9 (561,753) linear, 4 threads/pair>"),
        make_tile2_entry<T9R, kModeQuant>("tile2_soft_u8x2<K9 (561,753) ref-labels, 4 threads/pair>"),
        make_tile2_entry<T15, kModeQuant>("tile2_soft_u8x2<K15 (46321,51271) linear, 256 threads/pair>"),
    };
    *count = (int)(sizeof(e) / sizeof(e[0]));
    return e;
}

const Tile2Entry* tile2_find(uint32_t K, uint32_t n, uint32_t mode, const uint8_t* labels) {
    if (K < 8 || K > 15) return nullptr;
    int cnt = 0;
    const Tile2Entry* e = entries(&cnt);
    for (int i = 0; i < cnt; ++i)
        if (e[i].K == K && e[i].n == n && e[i].mode == mode && memcmp(e[i].labels, labels, 2u << (K - 1)) == 0) return &e[i];
    return nullptr;
}
const char* tile2_name(const Tile2Entry* e) { return e->name; }
int tile2_run(const Tile2Entry* e, const LaunchGeom& g, const void* in, uint8_t* out, int32_t* path_metric, uint32_t* scratch,
              cudaStream_t st) { return e->launch(g, in, out, path_metric, scratch, st); }
size_t tile2_bytes(const Tile2Entry* e, const LaunchGeom& g) { return e->bytes(g); }

}  // namespace vd
