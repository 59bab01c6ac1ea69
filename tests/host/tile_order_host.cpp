// Host build of the GEMM tile order (skyrl_b200/csrc/tile_order.cuh) for the CPU test suite.  TEST INFRASTRUCTURE ONLY.
#include "../../skyrl_b200/csrc/tile_order.cuh"

// Fills visits[m * n_tiles + n] with the number of times tile (m, n) is produced by t = 0 .. m_tiles * n_tiles - 1;
// returns the number of out-of-range coordinates.
extern "C" int skyb200_host_tile_order(int m_tiles, int n_tiles, int group, int order, int* visits, int* first_ms) {
    int bad = 0;
    for (int t = 0; t < m_tiles * n_tiles; ++t) {
        const skyb200::TileCoord c = skyb200::decode_tile(t, m_tiles, n_tiles, group, order);
        if (c.m_tile < 0 || c.m_tile >= m_tiles || c.n_tile < 0 || c.n_tile >= n_tiles) {
            ++bad;
            continue;
        }
        ++visits[c.m_tile * n_tiles + c.n_tile];
        if (first_ms != nullptr && t < 2 * m_tiles) first_ms[t] = (order == 0) ? c.m_tile : c.n_tile;
    }
    return bad;
}
