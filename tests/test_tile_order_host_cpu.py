"""The tile order of the persistent tcgen05 GEMM (skyrl_b200/csrc/tile_order.cuh, the source the kernel compiles)
built for the host with g++: for any tile counts and group size - including groups that do not divide the tile count,
groups larger than it, and the shapes of BASELINE configs 2-4 - both orders must visit every output tile exactly once,
and the fast index must sweep a whole group before the slow one advances (that is what makes concurrently resident
CTAs share operand panels through L2).  Test infrastructure only."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("host") / "tile_order_host.so"
    r = subprocess.run([gxx, "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC",
                        os.path.join(ROOT, "tests", "host", "tile_order_host.cpp"), "-o", str(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(str(so))
    ip = C.POINTER(C.c_int)
    lib.skyb200_host_tile_order.restype = C.c_int
    lib.skyb200_host_tile_order.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
    return lib


def _visits(lib, m, n, group, order):
    v = np.zeros(m * n, dtype=np.int32)
    first = np.full(2 * max(m, n) + 2, -1, dtype=np.int32)
    bad = lib.skyb200_host_tile_order(m, n, group, order, v.ctypes.data_as(C.POINTER(C.c_int)),
                                      first.ctypes.data_as(C.POINTER(C.c_int)) if order == 0 else None)
    return bad, v, first


def test_every_tile_exactly_once(host):
    rng = np.random.default_rng(0)
    cases = [(1, 1, 1), (1, 7, 4), (7, 1, 4), (5, 3, 8), (37, 594, 19), (37, 14, 8), (594, 14, 14),   # config 2
             (128, 594, 19), (128, 20, 8), (594, 20, 14), (64, 594, 19), (3, 594, 19)]               # configs 3, 4
    cases += [tuple(int(x) for x in rng.integers(1, 40, 3)) for _ in range(200)]
    for m, n, group in cases:
        for order in (0, 1):
            bad, v, _ = _visits(host, m, n, group, order)
            assert bad == 0 and (v == 1).all(), (m, n, group, order, int(bad), int((v != 1).sum()))


def test_fast_index_sweeps_the_group(host):
    # order 0, 37 M tiles in groups of 19: the first 19 work items are M tiles 0..18 of N tile 0, then N tile 1 ...
    _, _, first = _visits(host, 37, 594, 19, 0)
    assert first[:19].tolist() == list(range(19)) and first[19:38].tolist() == list(range(19))
