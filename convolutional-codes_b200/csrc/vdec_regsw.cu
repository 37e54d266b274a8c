// vdec_regsw.cu — registry of lane-group register kernels for K = 8 .. 12 (see vdec_regsw.cuh).
// Linear labels only; a request matches when its host-computed label table equals the entry's.
#include <string.h>

#include "vdec_regsw.cuh"

namespace vd {

using K8  = WideDef<8, 0xF900000000000000ull, 0xA700000000000000ull>;    // (371,247)_8
using K9  = WideDef<9, 0xB880000000000000ull, 0xF580000000000000ull>;    // (561,753)_8 — BASELINE config 4
using K10 = WideDef<10, 0x9DC0000000000000ull, 0xD940000000000000ull>;   // (1167,1545)_8
using K11 = WideDef<11, 0x9BA0000000000000ull, 0xF620000000000000ull>;   // (2335,3661)_8

static const WideEntry* entries(int* count) {
    static const WideEntry e[] = {
        make_wide_entry<K8>("regsw_soft_f32<K8 (371,247) linear, 2 lanes/frame>"),
        make_wide_entry<K9>("regsw_soft_f32<K9 (561,753) linear, 4 lanes/frame>"),
        make_wide_entry<K10>("regsw_soft_f32<K10 (1167,1545) linear, 8 lanes/frame>"),
        make_wide_entry<K11>("regsw_soft_f32<K11 (2335,3661) linear, 16 lanes/frame>"),
    };
    *count = (int)(sizeof(e) / sizeof(e[0]));
    return e;
}

const WideEntry* regsw_find(uint32_t K, uint32_t n, uint32_t metric, const uint8_t* labels) {
    if (metric != VDEC_SOFT_F32 || K < 8 || K > 12) return nullptr;
    int cnt = 0;
    const WideEntry* e = entries(&cnt);
    for (int i = 0; i < cnt; ++i)
        if (e[i].K == K && e[i].n == n && memcmp(e[i].labels, labels, 2u << (K - 1)) == 0) return &e[i];
    return nullptr;
}
const char* regsw_name(const WideEntry* e) { return e->name; }
int regsw_run(const WideEntry* e, const LaunchGeom& g, const void* in, uint8_t* out, uint32_t* scratch, cudaStream_t st) {
    return e->launch(g, in, out, nullptr, scratch, st);
}
size_t regsw_bytes(const WideEntry* e, const LaunchGeom& g) { return e->surv_bytes(g); }

}  // namespace vd
