// Tile order of the persistent GEMM (gemm_sm100.cuh): which (M tile, N tile) the t-th work item of the grid-stride
// loop is.  Pure integer code, so the same source also compiles for the host: tests/test_tile_order_host_cpu.py checks
// with g++ that every order visits every tile exactly once for any tile counts / group size (test infrastructure only).
#pragma once

#ifdef __CUDACC__
#define SKYB200_TILE_FN __device__ __forceinline__
#else
#include <algorithm>
#define SKYB200_TILE_FN inline
using std::min;
#endif

namespace skyb200 {

struct TileCoord {
    int m_tile;
    int n_tile;
};

// Tile order.  order 0: groups of `group` M-tiles sweep all N-tiles with the M index fastest, so concurrently
// resident CTAs share the same B tile (L2 hits) while the A group stays L2-resident.  order 1: the transpose -
// groups of `group` N-tiles sweep all M-tiles with the N index fastest (the B panels of the group stay hot while
// the A panels stream through once per group); used when N is small and K is short (d_weight GEMM).
SKYB200_TILE_FN TileCoord decode_tile(int t, int num_m_tiles, int num_n_tiles, int group, int order) {
    TileCoord c;
    if (order == 0) {
        const int tiles_per_group = group * num_n_tiles;
        const int g = t / tiles_per_group;
        const int first_m = g * group;
        const int gm = min(group, num_m_tiles - first_m);
        const int r = t - g * tiles_per_group;
        c.n_tile = r / gm;
        c.m_tile = first_m + (r - c.n_tile * gm);
    } else {
        const int tiles_per_group = group * num_m_tiles;
        const int g = t / tiles_per_group;
        const int first_n = g * group;
        const int gn = min(group, num_n_tiles - first_n);
        const int r = t - g * tiles_per_group;
        c.m_tile = r / gn;
        c.n_tile = first_n + (r - c.m_tile * gn);
    }
    return c;
}

}  // namespace skyb200
