// vdec_tile.cu — registry of register-tile kernels for K = 8 .. 15 (see vdec_tile.cuh).
// True-parity labels and the reference's de-facto labels; a request matches when its host-computed label table
// equals the entry's.
#include <string.h>

#include "vdec_tile.cuh"

namespace vd {

using T8  = TileDef<8,  0xF900000000000000ull, 0xA700000000000000ull>;    // (371,247)_8: the largest K the reference can run
using T8R = TileDef<8,  0xF900000000000000ull, 0xA700000000000000ull, true>;
using T9  = TileDef<9,  0xB880000000000000ull, 0xF580000000000000ull>;    // (561,753)_8 — BASELINE config 4
using T10 = TileDef<10, 0x9DC0000000000000ull, 0xD940000000000000ull>;    // (1167,1545)_8
using T11 = TileDef<11, 0x9BA0000000000000ull, 0xF620000000000000ull>;    // (2335,3661)_8
using T12 = TileDef<12, 0x8DD0000000000000ull, 0xBD30000000000000ull>;    // (4335,5723)_8
using T13 = TileDef<13, 0x8AD8000000000000ull, 0xFD88000000000000ull>;    // (10533,17661)_8
using T14 = TileDef<14, 0x8EF4000000000000ull, 0xB94C000000000000ull>;    // (21675,27123)_8
using T15 = TileDef<15, 0x99A2000000000000ull, 0xA572000000000000ull>;    // (46321,51271)_8 — BASELINE config 4

// the reference's de-facto labelling of the same polynomials (they carry taps at register positions 4 mod 8)
using T9R  = TileDef<9,  0xB880000000000000ull, 0xF580000000000000ull, true>;
using T11R = TileDef<11, 0x9BA0000000000000ull, 0xF620000000000000ull, true>;
using T13R = TileDef<13, 0x8AD8000000000000ull, 0xFD88000000000000ull, true>;
using T15R = TileDef<15, 0x99A2000000000000ull, 0xA572000000000000ull, true>;

// rate 1/3 on a large trellis: K = 9 (557,663,711)_8, true-parity labels
using T9N3 = TileDef<9, 0xB780000000000000ull, 0xD980000000000000ull, false, VDEC_TILE_VB, 3, 0xE480000000000000ull>;

static const TileEntry* entries(int* count) {
    static const TileEntry e[] = {
        make_tile_entry<T8>("tile_soft_f32<K8 (371,247) linear, 2 threads/frame>"),
        make_tile_entry<T8R>("tile_soft_f32<K8 (371,247) ref-labels, 2 threads/frame>"),
        make_tile_entry<T9>("tile_soft_f32<K9 (561,753) linear, 4 threads/frame>"),
        make_tile_entry<T10>("tile_soft_f32<K10 (1167,1545) linear, 8 threads/frame>"),
        make_tile_entry<T11>("tile_soft_f32<K11 (2335,3661) linear, 16 threads/frame>"),
        make_tile_entry<T12>("tile_soft_f32<K12 (4335,5723) linear, 32 threads/frame>"),
        make_tile_entry<T13>("tile_soft_f32<K13 (10533,17661) linear, 64 threads/frame>"),
        make_tile_entry<T14>("tile_soft_f32<K14 (21675,27123) linear, 128 threads/frame>"),
        make_tile_entry<T15>("tile_soft_f32<K15 (46321,51271) linear, 256 threads/frame>"),
        make_tile_entry<T9R>("tile_soft_f32<K9 (561,753) ref-labels, 4 threads/frame>"),
        make_tile_entry<T11R>("tile_soft_f32<K11 (2335,3661) ref-labels, 16 threads/frame>"),
        make_tile_entry<T13R>("tile_soft_f32<K13 (10533,17661) ref-labels, 64 threads/frame>"),
        make_tile_entry<T15R>("tile_soft_f32<K15 (46321,51271) ref-labels, 256 threads/frame>"),
        make_tile_entry<T9N3>("tile_soft_f32<K9 r1/3 (557,663,711) linear, 4 threads/frame>"),
    };
    *count = (int)(sizeof(e) / sizeof(e[0]));
    return e;
}

const TileEntry* tile_find(uint32_t K, uint32_t n, uint32_t metric, const uint8_t* labels) {
    if (metric != VDEC_SOFT_F32 || K < 8 || K > 15) return nullptr;
    int cnt = 0;
    const TileEntry* e = entries(&cnt);
    for (int i = 0; i < cnt; ++i)
        if (e[i].K == K && e[i].n == n && memcmp(e[i].labels, labels, 2u << (K - 1)) == 0) return &e[i];
    return nullptr;
}
const char* tile_name(const TileEntry* e) { return e->name; }
int tile_run(const TileEntry* e, const LaunchGeom& g, const void* in, uint8_t* out, uint32_t* scratch, cudaStream_t st) {
    return e->launch(g, in, out, scratch, st);
}
size_t tile_bytes(const TileEntry* e, const LaunchGeom& g) { return e->bytes(g); }

}  // namespace vd
