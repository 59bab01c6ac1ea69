"""GPU probe of the tcgen05 GEMM mainloop (all operand-major / cta_group variants) against torch.matmul.

Each configuration runs in its own subprocess under a timeout so a protocol bug cannot hang the box.
    python scripts/gemm_probe.py            # all configurations
    python scripts/gemm_probe.py one <ctas> <bn> <a_mn> <b_mn> <M> <N> <K>
"""
import json
import subprocess
import sys


def run_one(ctas, bn, a_mn, b_mn, M, N, K):
    import torch

    from skyrl_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device=dev) * 0.5).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    A_in = A.t().contiguous() if a_mn else A
    B_in = B.t().contiguous() if b_mn else B
    Cm = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
    rc = lib.skyb200_gemm_selftest(_lib.ptr(A_in), _lib.ptr(B_in), _lib.ptr(Cm), M, N, K, a_mn, b_mn, ctas, bn, 0,
                                   _lib.stream_ptr())
    out = {"cfg": [ctas, bn, a_mn, b_mn, M, N, K], "rc": rc}
    if rc != 0:
        out["err"] = lib.skyb200_last_error().decode()
        print(json.dumps(out))
        return 1
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        out["sync_error"] = str(e)[:300]
        out["dbg"] = _lib.gemm_debug_record()
        print(json.dumps(out))
        return 2
    err = (Cm - ref).abs()
    out["max_err"] = float(err.max())
    out["nan"] = int(torch.isnan(Cm).sum())
    out["ref_absmax"] = float(ref.abs().max())
    if out["nan"] or out["max_err"] > 1e-2 * max(1.0, out["ref_absmax"]):
        bad = (err > 1e-2 * max(1.0, out["ref_absmax"])) | torch.isnan(Cm)
        idx = bad.nonzero()
        out["num_bad"] = int(bad.sum())
        out["first_bad"] = idx[:4].tolist()
        out["bad_rows"] = sorted(set((idx[:, 0] // 8).tolist()))[:16]
        out["bad_cols"] = sorted(set((idx[:, 1] // 8).tolist()))[:16]
    # timing of a bigger run for a first throughput number
    print(json.dumps(out))
    return 0 if ("num_bad" not in out) else 3


CONFIGS = [
    # ctas, bn, a_mn, b_mn, M, N, K
    (1, 128, 0, 0, 128, 128, 64),
    (1, 128, 0, 0, 256, 384, 256),
    (1, 256, 0, 0, 384, 512, 320),
    (1, 256, 0, 0, 200, 300, 136),   # ragged tails (TMA zero fill, predicated stores)
    (2, 256, 0, 0, 256, 256, 64),
    (2, 256, 0, 0, 768, 1024, 512),
    (2, 256, 0, 0, 1000, 1288, 328),
    (2, 128, 0, 0, 512, 384, 256),
    (1, 256, 0, 1, 256, 512, 256),   # B MN-major
    (1, 256, 1, 0, 256, 512, 256),   # A MN-major
    (1, 256, 1, 1, 256, 512, 256),
    (2, 256, 0, 1, 512, 512, 256),
    (2, 256, 1, 1, 512, 768, 320),
    (2, 256, 1, 1, 1000, 904, 200),
    (2, 256, 0, 0, 9472, 4096, 3584),  # many tiles per CTA, deep K: pipeline wrap-around
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        sys.exit(run_one(*[int(x) for x in sys.argv[2:9]]))
    fails = 0
    for cfg in CONFIGS:
        cmd = [sys.executable, __file__, "one", *map(str, cfg)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            print(r.stdout.strip() or f"NO OUTPUT cfg={cfg} rc={r.returncode}")
            if r.returncode != 0:
                fails += 1
                print("  stderr:", r.stderr.strip()[-600:])
        except subprocess.TimeoutExpired:
            fails += 1
            print(f"TIMEOUT cfg={cfg}")
    print("FAILS", fails)
    sys.exit(1 if fails else 0)
