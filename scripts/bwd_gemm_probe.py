"""Sustained-regime rate of the three GEMMs of the fused step, each alone and back to back for ~1.5 s (power-capped
clocks), on controlled operand data - to separate what the kernel FORM costs (operand majors, epilogue, tile order)
from what the DATA costs (a power-capped part slows down on operands that toggle more bits) and from what running
next to the auxiliary-stream kernels costs inside the step.

    python scripts/bwd_gemm_probe.py [rows V H]

dS variants: "randn" = N(0, 1) bf16 (what a GEMM benchmark uses); "dlogits" = realistic d(logits) of a policy step:
g * (onehot - softmax(z)) with z ~ N(0, 1.5^2), g ~ N(0, 1) / N_tokens (tiny values, wide exponent range);
"zeros" = the floor (no operand toggling on one side).
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import ClockSampler  # noqa: E402
from skyrl_b200 import _lib  # noqa: E402


def sustained(fn, seconds=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    cs = ClockSampler(0)
    cs.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 0
    s.record()
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        for _ in range(10):
            fn()
        reps += 10
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    c = cs.stop()
    return s.elapsed_time(e) / reps, c.get("sm_mhz")


def main():
    kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    pos = [a for a in sys.argv[1:] if "=" not in a]
    rows, V, H = (int(x) for x in pos[:3]) if len(pos) >= 3 else (9472, 152064, 3584)
    quick = kv.get("quick") == "1"  # A/B mode: realistic d(logits) only, no cuBLAS legs
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    hidden = torch.randn(rows, H, device=dev, generator=g).bfloat16()
    weight = (torch.randn(V, H, device=dev, generator=g) / H ** 0.5).bfloat16()
    ldS = (V + 63) // 64 * 64
    S = torch.zeros(rows, ldS, dtype=torch.bfloat16, device=dev)
    d_hidden = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
    d_weight = torch.zeros(V, H, dtype=torch.float32, device=dev)
    st = _lib.stream_ptr()
    flops = 2.0 * rows * V * H
    P = _lib.ptr

    def fill(kind):
        if kind == "zeros":
            S.zero_()
        elif kind == "randn":
            for r0 in range(0, rows, 1024):
                S[r0:r0 + 1024, :V] = torch.randn(min(1024, rows - r0), V, device=dev, generator=g)
        else:
            for r0 in range(0, rows, 1024):
                n = min(1024, rows - r0)
                z = (hidden[r0:r0 + n].float() @ weight.float().t()) * 1.5
                p = torch.softmax(z, -1)
                lab = torch.randint(0, V, (n,), device=dev, generator=g)
                p[torch.arange(n, device=dev), lab] -= 1.0
                gl = torch.randn(n, 1, device=dev, generator=g) / 1048576.0
                S[r0:r0 + n, :V] = (-gl * p)

    out = {}
    for kind in (("dlogits",) if quick else ("randn", "dlogits", "zeros")):
        fill(kind)
        for name, call in (
            ("d_hidden", lambda: _lib.check(lib.skyb200_head_gemm_bench(1, None, P(weight), P(S), P(d_hidden), None, 0,
                                                                        rows, H, V, st))),
            ("d_weight_acc", lambda: _lib.check(lib.skyb200_head_gemm_bench(2, P(hidden), None, P(S), None, P(d_weight),
                                                                            1, rows, H, V, st))),
            ("d_weight_store", lambda: _lib.check(lib.skyb200_head_gemm_bench(2, P(hidden), None, P(S), None,
                                                                              P(d_weight), 0, rows, H, V, st))),
        ):
            d_weight.zero_()
            ms, mhz = sustained(call)
            out[f"{name}[{kind}]"] = (round(flops / ms / 1e9, 1), round(ms, 3), mhz)
        if kind == "dlogits":  # full-size spot check of the form that just ran (the last call was a plain store)
            rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-300))
            out["err_d_weight_rows"] = rel(d_weight[-300:], S[:, V - 300:V].float().t() @ hidden.float())
            _lib.check(lib.skyb200_head_gemm_bench(1, None, P(weight), P(S), P(d_hidden), None, 0, rows, H, V, st))
            out["err_d_hidden_rows"] = rel(d_hidden[-300:].float(), S[-300:, :V].float() @ weight.float())
    if quick:
        print(json.dumps({"tag": kv.get("tag", ""), "tflops_ms_mhz": out}), flush=True)
        return
    a = hidden
    c = torch.empty(rows, V, dtype=torch.bfloat16, device=dev)
    ms, mhz = sustained(lambda: torch.matmul(a, weight.t(), out=c))
    out["cublas_fwd_shape"] = (round(flops / ms / 1e9, 1), round(ms, 3), mhz)
    Sv = S[:, :V]
    fill("dlogits")
    dh32 = torch.empty(rows, H, dtype=torch.bfloat16, device=dev)
    ms, mhz = sustained(lambda: torch.matmul(Sv, weight, out=dh32))
    out["cublas_d_hidden[dlogits]"] = (round(flops / ms / 1e9, 1), round(ms, 3), mhz)
    dwb = torch.empty(V, H, dtype=torch.bfloat16, device=dev)
    ms, mhz = sustained(lambda: torch.matmul(Sv.t(), hidden, out=dwb))
    out["cublas_d_weight_bf16out[dlogits]"] = (round(flops / ms / 1e9, 1), round(ms, 3), mhz)
    print(json.dumps({"shape": {"rows": rows, "V": V, "H": H}, "tflops_ms_mhz": out}))


if __name__ == "__main__":
    main()
