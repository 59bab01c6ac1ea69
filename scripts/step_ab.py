"""Time the fused policy step (config 2, device-resident inputs) for the library / env of THIS process.
A/B runs alternate processes inside one gpurun call (box-to-box variance is larger than most effects):

    for i in 1 2; do for v in 0 296; do SKYB200_TRANSFORM_BLOCKS=$v python scripts/step_ab.py tag=$v; done; done
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import skyrl_b200  # noqa: E402
from bench import ClockSampler, make_cfg  # noqa: E402


def main():
    kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    steps, warmup = int(kv.get("steps", 4)), int(kv.get("warmup", 2))
    B, A, H, V = 256, 4096, 3584, 152064
    N = B * A
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1)
    hidden = torch.empty((N, H), dtype=torch.bfloat16, device=dev)
    for r0 in range(0, N, 65536):
        hidden[r0:r0 + 65536] = torch.randn((min(65536, N - r0), H), generator=gen, device=dev)
    weight = (torch.randn((V, H), generator=gen, device=dev) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=gen, device=dev)
    mask = torch.ones(B, A, device=dev)
    if kv.get("ragged"):  # right-padded responses: lengths uniform in [A/8, A] (mean valid fraction ~56 %)
        lens = torch.randint(A // 8, A + 1, (B,), generator=gen, device=dev)
        mask = (torch.arange(A, device=dev)[None] < lens[:, None]).float()
    adv = torch.randn((B, 1), generator=gen, device=dev).expand(B, A).contiguous()
    cfg = make_cfg()  # regular / token_mean / k3 KL loss: the bench configuration
    if kv.get("data", "bench") == "bench":
        # the data of bench.py: old / base log-probs next to the current policy's (near on-policy), so that nearly
        # every token carries a gradient - operand data matters on a power-capped part
        with torch.no_grad():
            lp0 = skyrl_b200.fused_lm_head_logprobs(hidden, weight, labels.view(N)).view(B, A)
        olp = lp0 + 0.05 * torch.randn((B, A), generator=gen, device=dev)
        base = lp0 + 0.1 * torch.randn((B, A), generator=gen, device=dev)
        del lp0
    else:  # round-1 data: far-off-policy, half of the sequences fully clipped (zero gradient rows)
        olp = -torch.rand((B, A), generator=gen, device=dev) * 12
        base = olp
    dw = torch.empty((V, H), dtype=torch.float32, device=dev)
    timing = kv.get("timing", "1") == "1" and hasattr(skyrl_b200, "step_gemm_times")

    def step():
        kw = dict(timing=True) if timing else {}
        return skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask, base_log_probs=base,
                                             d_weight=dw, sync=False, **kw)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if timing:
        skyrl_b200.step_gemm_times(dev)  # drop the warm-up record
    cs = ClockSampler(0)
    cs.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step()
    e.record()
    torch.cuda.synchronize()
    clocks = cs.stop()
    ms = s.elapsed_time(e) / steps
    gemm = None
    if timing:
        gemm = {k: round(v["ms"] / max(v["launches"], 1), 3) for k, v in skyrl_b200.step_gemm_times(dev).items()}
    print(json.dumps({"tag": kv.get("tag", ""), "ms_per_step": round(ms, 1), "tok_s": round(N / ms * 1e3), "gemm_ms": gemm,
                      "valid_frac": round(float(mask.mean()), 3), "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}), flush=True)


if __name__ == "__main__":
    main()
