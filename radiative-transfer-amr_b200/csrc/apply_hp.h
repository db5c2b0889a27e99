// Group sizing shared by the hp apply kernel (apply_hp.cu) and the bucket builder (apply.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

constexpr int HP_THREADS = 256;
constexpr int HP_MT = 8;          // m-tiles per block (R <= 64 rows)
constexpr int HP_MAXCG = 7;
constexpr int HP_ROWS = HP_MT * 8;

__host__ __device__ inline int hp_ntps(int ntp) {
    int v = (ntp + 7) & ~7;       // multiple of 8, = 8 (mod 16): conflict-free A fragments
    while ((v & 15) != 8) v += 8;
    return v;
}
// shared memory (doubles) of a group: U | FA (cg x 2 faces x n x ntp) | TH (wt, ct, st, wF0..3) | bits | mk | ms | y offsets | row info
__host__ __device__ inline size_t hp_smem_doubles(int n, int cg, int nt) {
    const int ntp = nt + (nt & 1), k8 = (nt + 7) & ~7, rp = (cg * n * n + 7) & ~7;
    return (size_t)rp * hp_ntps(ntp) + (size_t)cg * 2 * n * ntp + 7 * (size_t)k8 + k8 / 2 +
           (size_t)rp * 3 + rp / 2;
}
// One record per group (thread block): everything the block needs to start its loads after a
// single global round trip.
struct HpGroup {
    int32_t cg_n, sg, nt, ntp;
    int64_t tp0, g_off;
    int32_t cols[8];
    int64_t slab_off[8];
};
inline bool hp_supported_n(int n) { return n >= 2 && n <= 8; }
