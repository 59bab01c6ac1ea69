"""Context measurement: the REFERENCE's own GPU path for the same work, on the same B200.

What skyrl_train runs per micro-batch (model_wrapper.py:309-363, worker.py:787-877), written with the same
library calls it uses: bf16 F.linear (cuBLASLt) -> logits.div_(T) -> flash-attn Triton cross-entropy
(inplace_backward=True) -> chunked entropy (1024-position chunks, log_softmax/exp/sum) -> policy loss + KL term
(torch elementwise) -> backward.  It materialises [tokens, vocab] logits, so it only fits micro-batches; tokens/s
is reported for the micro-batch size given (default 8192 tokens, the largest that is comfortable in fp32 temps).

This is NOT the oracle and not part of the product, and it imports neither: the few torch elementwise ops of the
loss are written out below (/root/reference is absent on the GPU box).  Output: one JSON line.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F


def main():
    tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    H, V, A = 3584, 152064, 4096
    dev = torch.device("cuda:0")
    from flash_attn.ops.triton.cross_entropy import cross_entropy_loss

    g = torch.Generator(device=dev).manual_seed(0)
    hidden = torch.randn(tokens, H, device=dev, generator=g).bfloat16().requires_grad_(True)
    weight = (torch.randn(V, H, device=dev, generator=g) / H ** 0.5).bfloat16().requires_grad_(True)
    labels = torch.randint(0, V, (tokens,), device=dev, generator=g)
    rows = max(1, tokens // A)
    a_len = tokens // rows
    olp = -12.0 + 0.05 * torch.randn(rows, a_len, device=dev, generator=g)
    base = -12.0 + 0.1 * torch.randn(rows, a_len, device=dev, generator=g)
    adv = torch.randn(rows, a_len, device=dev, generator=g)
    mask = torch.ones(rows, a_len, device=dev)
    eps, kl_coef = 0.2, 1e-3

    def step():
        hidden.grad = None
        weight.grad = None
        logits = F.linear(hidden, weight)                     # HF lm_head, bf16
        logits.div_(1.0)                                      # model_wrapper.py:314
        lp = -cross_entropy_loss(logits, labels, inplace_backward=True)[0]   # torch_utils.py:148-153
        with torch.no_grad():                                 # use_entropy_loss=False: metric only
            ent = torch.zeros(tokens, dtype=logits.dtype, device=dev)
            for s in range(0, tokens, 1024):                  # torch_utils.py:59-111
                lsm = F.log_softmax(logits[s:s + 1024], dim=-1)
                ent[s:s + 1024] = -(lsm.exp() * lsm).sum(-1)
        lp2 = lp.view(rows, a_len)
        ratio = torch.exp(torch.clamp(lp2 - olp, -20.0, 20.0))            # ppo_utils.py:548-586, "regular"
        surr = -torch.min(ratio * adv, ratio.clamp(1 - eps, 1 + eps) * adv)
        denom = mask.sum().clamp(min=1.0)
        policy = (surr * mask).sum() / denom                               # token_mean
        with torch.no_grad():                                              # k3 KL term: value only (worker.py:865)
            x = torch.clamp(base - lp2, -20.0, 20.0)
            kl = (torch.clamp(torch.exp(x) - x - 1, -10.0, 10.0) * mask).sum(-1) / mask.sum(-1).clamp(min=1.0)
            ent_mean = (ent.view(rows, a_len).float() * mask).sum() / denom
        final = policy + kl.mean() * kl_coef - ent_mean * 0.0
        final.backward()
        return final

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    s.record()
    for _ in range(n):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    print(json.dumps({"what": "reference GPU path (F.linear bf16 + flash-attn CE + chunked entropy + torch loss), "
                              "Qwen2.5-7B head", "micro_batch_tokens": tokens, "ms_per_micro_batch": ms,
                      "tokens_per_s": tokens / (ms * 1e-3),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
