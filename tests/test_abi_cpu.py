"""The C-ABI shared library builds, loads on a CPU-only box and exports every symbol include/skyb200.h declares.
No compute call is made here (no GPU)."""
import ctypes
import os
import re

import pytest

from skyrl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "skyb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skyb200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not _lib.lib_path().exists():
        from skyrl_b200 import build

        build.build()
    return _lib.load()


def test_header_and_binding_agree():
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SIGNATURES) == declared


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(str(_lib.lib_path()))
    for name in _declared_symbols():
        assert hasattr(raw, name), f"{name} declared in skyb200.h but not exported by libskyb200.so"


def test_host_only_entry_points(lib):
    assert lib.skyb200_abi_version() == 2
    assert lib.skyb200_policy_loss_workspace(256, 4096) >= 256 * 8
    assert lib.skyb200_gae_workspace(512) >= 512 * 24
    fwd = lib.skyb200_head_forward_workspace(1 << 20, 3584, 152064, 0)
    bwd = lib.skyb200_head_backward_workspace(1 << 20, 3584, 152064, 0)
    # bounded scratch: independent of N once N exceeds one chunk, and far below N*V*2 bytes (= 319 GB)
    assert fwd == lib.skyb200_head_forward_workspace(1 << 22, 3584, 152064, 0)
    assert bwd == lib.skyb200_head_backward_workspace(1 << 22, 3584, 152064, 0)
    assert bwd < 7 * 2**30 and fwd < 2**30


def _plan(lib, B, A, H, V, loss, n_tokens=0, need_grad=1):
    out = (ctypes.c_int64 * 3)()
    assert lib.skyb200_policy_step_plan(B, A, n_tokens, H, V, _lib.LOSS_TYPES[loss], need_grad, out) == 0
    ws = lib.skyb200_policy_step_workspace(B, A, n_tokens, H, V, _lib.LOSS_TYPES[loss], need_grad)
    return out[0], out[1], out[2], ws


def test_fused_step_chunk_plan(lib):
    """Host-only planning of the fused step: bounded scratch for every loss and any padded response length."""
    H, V = 3584, 152064
    cap, stride, n, ws = _plan(lib, 256, 4096, H, V, "regular")
    assert (cap, stride, n) == (18944, 18944, 56) and ws < 13 * 2**30          # 74 pair-tiles, 2 x 5.8 GB store
    assert _plan(lib, 1024, 4096, H, V, "regular")[3] - ws < 2**20             # independent of the batch size
    assert _plan(lib, 4, 100, H, V, "dual_clip")[:3] == (512, 512, 1)          # a batch smaller than one chunk
    # GSPO: whole sequences per chunk; the chunk count covers the batch exactly
    for B, A in ((256, 4096), (8, 32768), (40, 1000), (256, 3917), (7, 4099), (3, 70000), (5, 1)):
        cap, stride, n, ws = _plan(lib, B, A, H, V, "gspo")
        assert stride % A == 0 and cap % 256 == 0 and stride <= cap < stride + 256, (B, A, cap, stride)
        assert n == -(-B * A // stride)
        assert cap <= max(18944 * 3 // 2, -(-A // 256) * 256), (B, A, cap)       # ~1.5 default chunks, or one sequence
        assert ws < 2 * cap * (V + 64) * 2 + 2 * cap * (3 * 595 + 6) * 4 + (B + n) * 64 + 2**22, (B, A, ws)
    # a response length that shares no factor with the 256-row tile used to need lcm(A, 256) = 1 M tokens (611 GB)
    assert _plan(lib, 256, 3917, H, V, "gspo")[3] < 18 * 2**30
    # the packed layout plans over n_tokens
    assert _plan(lib, 64, 1, H, V, "regular", n_tokens=500000)[2] == -(-500000 // 18944)


def test_d_weight_tile_groups_fit_l2(lib):
    """Host-only: the N-group size of the d_weight GEMM.  A group's hidden-chunk panels (group x rows x 256 bf16) stay
    under 36 MB whenever that is cheaper than re-reading them once per round of M tiles; groups are balanced."""
    V = 152064
    assert lib.skyb200_dw_group_plan(9472, 3584, V) == 7        # 7B head: two groups of 7 (34 MB each), not 14 (68 MB)
    assert lib.skyb200_dw_group_plan(9472, 896, 151936) == 4    # 0.5B head: all 4 N tiles fit (19 MB)
    assert lib.skyb200_dw_group_plan(9472, 5120, 151936) == 7   # 32B-shaped head: 7 / 7 / 6
    assert lib.skyb200_dw_group_plan(32768, 5120, 151936) == 10  # one 32768-token GSPO sequence: nothing fits, 2 x 10
    assert lib.skyb200_dw_group_plan(4352, 3584, V) == 14       # a short last chunk: the whole hidden chunk fits
    for rows in (1, 300, 5000, 9472, 18944, 70000):
        for H in (64, 256, 896, 3584, 5120, 8192):
            g = lib.skyb200_dw_group_plan(rows, H, V)
            n = -(-H // 256)
            groups = -(-n // g)
            assert 1 <= g <= n and -(-n // groups) == g, (rows, H, g)


def _tiles(lib, rows, H, V):
    out = (ctypes.c_int32 * 4)()
    assert lib.skyb200_head_tile_plan(rows, H, V, out) == 0
    return list(out)


def test_backward_gemm_tile_widths(lib):
    """Host-only: which GEMMs of a chunk run on the wide 256 x 512 tiles.  The forward never does (its epilogue needs the
    second accumulator); the backward GEMMs do unless tile quantisation on the 74 CTA pairs costs more than the 25 %
    of operand traffic the wide form saves."""
    V = 152064
    assert _tiles(lib, 18944, 3584, V) == [256, 512, 512, 7]    # the default chunk: 74 x 7 tiles = 7 whole waves
    assert _tiles(lib, 9472, 3584, V)[:3] == [256, 512, 512]    # 37 x 7 = 259 tiles = 3.5 waves: still ahead (measured)
    assert _tiles(lib, 6656, 3584, V)[1] == 256                 # config 2's last chunk: 182 tiles would idle 18 % of 3 waves
    assert _tiles(lib, 18944, 896, 151936)[:3] == [256, 512, 512]  # 0.5B head: 896 = 1.75 wide tiles = 3.5 narrow ones
    assert _tiles(lib, 18944, 256, V)[1:3] == [256, 256]        # half of every wide tile would be padding
    assert _tiles(lib, 32768, 5120, 151936)[:3] == [256, 512, 512]  # one GSPO sequence on the 32B-shaped head
    for rows in (1, 300, 5000, 9472, 18944, 70000):
        for H in (64, 256, 896, 3584, 5120, 8192):
            t = _tiles(lib, rows, H, V)
            n = -(-H // t[2])
            assert t[0] == 256 and t[1] in (256, 512) and t[2] in (256, 512) and 1 <= t[3] <= n, (rows, H, t)


def test_loss_params_layout_matches_header():
    assert ctypes.sizeof(_lib.LossParams) == 6 * 4 + 8 * 4


def test_bad_arguments_report_errors_without_a_gpu(lib):
    # argument validation happens before any CUDA call
    p = _lib.LossParams()
    rc = lib.skyb200_policy_loss(None, None, None, None, None, None, None, None, None, None, 0, 4, 4,
                                 ctypes.byref(p), None, None, None)
    assert rc == 1
    assert b"null pointer" in lib.skyb200_last_error()
    rc = lib.skyb200_approx_kl(ctypes.c_void_p(16), ctypes.c_void_p(16), None, ctypes.c_void_p(16), 4, 9, None)
    assert rc == 1 and b"Invalid KL estimator type" in lib.skyb200_last_error()


def test_header_is_plain_c_and_a_c_program_links_the_library(lib, tmp_path):
    """The boundary is a C ABI: the header must compile as C11 (no C++-isms), and a C program with no Python or
    torch in sight links libskyb200.so and calls it (host-only entry points: runs without a GPU)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(inc, "skyb200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    libdir = str(_lib.lib_path().parent)
    exe = str(tmp_path / "c_abi_client")
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", inc,
                        os.path.join(root, "tests", "host", "c_abi_client.c"), "-L", libdir,
                        "-l:" + _lib.lib_path().name, "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.split(maxsplit=5)
    assert out[:3] == ["18944", "18944", "56"] and int(out[3]) > 10 * 2**30 and out[4] == "1"
    assert "null pointer" in out[5]

