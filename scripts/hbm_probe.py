"""Time the memory-bound kernels of the path alone (the `hbm_kernels` section of bench.py), plus calibration points
measured the same way (L2 flushed, CUDA events, median of 15): a torch copy of the same size and a tiny call."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from skyrl_b200 import _lib  # noqa: E402


def timed(fn, flush, reps=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    res = bench.hbm_kernels(dev, bench.measured_peaks())
    for k, v in res.items():
        print(k, json.dumps(v) if isinstance(v, dict) else v)
    if "quick" not in sys.argv:
        for k, v in bench.hbm_kernels(dev, bench.measured_peaks(), B=2048).items():
            print("4x", k, json.dumps(v) if isinstance(v, dict) else v)
    if "quick" in sys.argv:
        sys.exit(0)
    flush = torch.zeros(96 << 20, dtype=torch.int32, device=dev)
    B, A = 512, 8192
    x = torch.randn(B, A, device=dev)
    y = torch.empty_like(x)
    us = timed(lambda: y.copy_(x), flush)
    print("calib torch copy 16.8 MB -> 16.8 MB: %.1f us = %.0f GB/s" % (us, 2 * x.numel() * 4 / us / 1e3))
    xs = [torch.randn(B, A, device=dev) for _ in range(5)]
    us = timed(lambda: torch.stack(xs).sum(0, out=y) if False else torch.add(xs[0], xs[1], out=y), flush)
    print("calib torch add (2 reads + 1 write, 50 MB): %.1f us = %.0f GB/s" % (us, 3 * x.numel() * 4 / us / 1e3))
    big = torch.randn(8 * B, A, device=dev)
    big2 = torch.empty_like(big)
    us = timed(lambda: big2.copy_(big), flush)
    print("calib torch copy 134 MB -> 134 MB: %.1f us = %.0f GB/s" % (us, 2 * big.numel() * 4 / us / 1e3))
    us = timed(lambda: None, flush)
    print("calib empty region: %.1f us" % us)
    lib = _lib.load()
    # GAE on few long rows (config-4 shard: 8 trajectories x 32768 positions): rows are split into one block per tile
    for (b, a) in ((8, 32768), (64, 8192), (512, 8192)):
        r, v = torch.randn(b, a, device=dev), torch.randn(b, a, device=dev)
        m = torch.ones(b, a, device=dev)
        o1, o2 = torch.empty_like(r), torch.empty_like(r)
        st4 = torch.zeros(4, device=dev)
        ws = torch.empty(lib.skyb200_gae_workspace(b), dtype=torch.uint8, device=dev)
        call = lambda: _lib.check(lib.skyb200_gae_advantage(_lib.ptr(r), _lib.ptr(v), _lib.ptr(m), _lib.MASK_F32,
                                                            _lib.ptr(o1), _lib.ptr(o2), _lib.ptr(st4), _lib.ptr(ws),
                                                            ws.numel(), b, a, 1.0, 0.95, _lib.stream_ptr()))
        us = timed(call, flush)
        print("gae %d x %d: %.1f us = %.0f GB/s (28 B/position)" % (b, a, us, b * a * 28 / us / 1e3))
    t = torch.zeros(8, 64, device=dev)
    o = torch.zeros(8, 64, device=dev)
    us = timed(lambda: lib.skyb200_approx_kl(_lib.ptr(t), _lib.ptr(t), None, _lib.ptr(o), 512, 3, _lib.stream_ptr()), flush)
    print("calib one tiny kernel through the C ABI: %.1f us" % us)
