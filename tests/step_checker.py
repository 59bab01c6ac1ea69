"""Checker for one fused policy-update step at sizes where the CPU oracle cannot evaluate the whole batch
(BASELINE configs 2-5).  Device-agnostic test infrastructure: everything is pulled to the CPU and compared with
`oracle/` there.

What is checked, for any loss / reduction / layout:

* status dict: the oracle's loss assembly (worker.py:852-889) evaluated on the step's OWN log-probs and entropies
  (elementwise fp32 work over [B, A], cheap at any size);
* sampled valid token rows: log-prob and entropy against the fp32-logit oracle head on those rows only, and
  d_hidden against the oracle head's backward driven by the oracle's d(final)/d(log-prob) and d(final)/d(entropy)
  at those tokens (so the loss gradient, the 1/T, the entropy-gradient term and the d_hidden GEMM are all covered);
* sampled d_weight rows (vocabulary entries, half of them label ids so the one-hot term is exercised) against an fp32
  reference built on the step's device from the definition: d_weight[v, :] = sum_t d_logits[t, v] * hidden[t, :] with
  d_logits from fp32 logits of those columns only (N x 16 instead of N x V) - an ORACLE VALUE for d_weight at full size;
* size-independent identities over the whole result: sum_v d_weight[v, :] = 0 (every d_logits row sums to zero
  over the vocabulary, with or without the entropy term) and <d_weight, weight> = <d_hidden, hidden> (both are
  sum_{t,v} d_logits[t, v] * logits[t, v]) - a checksum of checksums that one lost output tile breaks.
"""
import torch

from oracle import policy_oracle as O


def d_weight_rows_reference(hidden, weight, labels, log_probs, entropy, g_lp, g_ent, T, rows, chunk=32768):
    """fp32 reference of d_weight[rows, :] on hidden's device.  All per-token arrays are flat [N] in the step's token
    layout; log_probs / entropy are the step's own forward outputs (checked separately), g_lp / g_ent the oracle's
    d(final)/d(log-prob), d(final)/d(entropy).  d_logits[t, v] = (g_lp * (1[v == label] - p) - g_ent * p * (log p + H)) / T
    with p = exp(z[t, v] / T - lse[t]), lse[t] = z[t, label] / T - log_prob[t]."""
    dev = hidden.device
    rows = rows.to(dev)
    w_rows = weight[rows].float()                                     # [R, H]
    out = torch.zeros(rows.numel(), hidden.shape[1], dtype=torch.float64, device=dev)
    N = hidden.shape[0]
    for t0 in range(0, N, chunk):
        sl = slice(t0, min(N, t0 + chunk))
        h = hidden[sl].float()
        lab = labels[sl].to(dev)
        z_lab = torch.cat([(h[i0:i0 + 8192] * weight[lab[i0:i0 + 8192]].float()).sum(-1)
                           for i0 in range(0, h.shape[0], 8192)])
        lse = z_lab / T - log_probs[sl].to(dev).float()
        logp = (h @ w_rows.t()) / T - lse[:, None]                    # [n, R] log-probabilities of the sampled columns
        p = logp.exp()
        glp, gent, H_t = g_lp[sl].to(dev).float()[:, None], g_ent[sl].to(dev).float()[:, None], entropy[sl].to(dev).float()[:, None]
        onehot = (lab[:, None] == rows[None, :]).float()
        d = (glp * (onehot - p) - gent * p * (logp + H_t)) / T
        out += (d.t().double() @ h.double())
    return out


def check_step(out, hidden, weight, labels, olp, adv, cfg, mask, base, T, keep=None, n_sample=32, seed=0,
               lp_tol=1e-3, grad_rel=2e-2, status_rel=1e-5):
    """`out`: PolicyUpdateOutput.  `hidden` [N, H] bf16 and `labels` [N] in the step's token layout (N = B*A padded,
    or the packed count); `olp`, `adv`, `mask`, `base` are PADDED [B, A] tensors; `keep` ([B, A] bool, packed layout
    only) marks the padded positions that exist in the token layout.  Returns a dict of the measured errors."""
    cpu = lambda t: None if t is None else t.detach().cpu()
    olp, adv, mask, base = cpu(olp).float(), cpu(adv).float(), cpu(mask).float(), cpu(base)
    B, A = olp.shape
    if keep is None:
        to_pad = lambda flat: cpu(flat).reshape(B, A)
        flat_index = torch.arange(B * A).reshape(B, A)
    else:
        keep = cpu(keep).bool()

        def to_pad(flat):
            x = torch.zeros(B, A, dtype=torch.float32)
            x[keep] = cpu(flat).float().reshape(-1)
            return x

        flat_index = torch.full((B, A), -1, dtype=torch.int64)
        flat_index[keep] = torch.arange(int(keep.sum()))
    res = {}

    # ---- status + the oracle's gradients w.r.t. log-prob / entropy, on the step's own forward outputs
    lp = to_pad(out.log_probs).clone().requires_grad_(True)
    ent = to_pad(out.entropy).clone().requires_grad_(True)
    final, st = O.loss_assembly(lp, olp, adv, cfg, mask, base.float() if base is not None else None, ent)
    final.backward()
    keys = ["final_loss", "policy_loss", "policy_entropy", "loss_metrics/clip_ratio"]
    if cfg.use_kl_loss and base is not None:
        keys.append("policy_kl")
    # north_star: loss within 1e-5 relative in fp32.  The policy loss is a sum of signed terms that largely cancel
    # (advantages are zero-mean), and an fp32 sum - the reference's as much as ours - is accurate relative to the L1
    # mass of its terms, not to the cancelled result: the loss keys are held to 1e-5 of max(|value|, L1 mass), where
    # the L1 mass is the same loss with |advantages| (every term then has one sign).  All-positive metrics (entropy,
    # KL, clip ratio) are held to plain 1e-5 relative.
    with torch.no_grad():
        l1 = abs(float(O.loss_assembly(lp.detach(), olp, adv.abs(), cfg, mask, None, None)[1]["policy_loss"])) \
            if cfg.policy_loss_type in ("regular", "dual_clip", "gspo") else 0.0
    for k in keys:
        got, want = out.status[k], st[k]
        scale = max(abs(want), l1) if k in ("final_loss", "policy_loss") else abs(want)
        assert abs(got - want) <= status_rel * scale + 1e-9, (k, got, want, scale)
    g_lp = lp.grad
    g_ent = ent.grad if ent.grad is not None else torch.zeros(B, A)

    # ---- sampled valid rows against the oracle head (forward and backward)
    valid = torch.nonzero((mask > 0) & (flat_index >= 0))
    assert valid.shape[0] > 0, "no valid token to sample"
    perm = torch.randperm(valid.shape[0], generator=torch.Generator().manual_seed(seed))[:n_sample]
    pb, pa = valid[perm, 0], valid[perm, 1]
    pick = flat_index[pb, pa]
    dev_pick = pick.to(hidden.device)
    hs = hidden[dev_pick].float().cpu().requires_grad_(True)
    w32 = weight.float().cpu()
    lp_s, ent_s = O.head_logprobs_entropy(hs, w32, labels.reshape(-1)[dev_pick].cpu(), T)
    res["lp_err"] = float((cpu(out.log_probs).reshape(-1)[pick] - lp_s.detach()).abs().max())
    res["ent_err"] = float((cpu(out.entropy).reshape(-1)[pick] - ent_s.detach()).abs().max())
    assert res["lp_err"] < lp_tol and res["ent_err"] < lp_tol, res
    (lp_s * g_lp[pb, pa] + ent_s * g_ent[pb, pa]).sum().backward()
    assert float(hs.grad.norm()) > 0.0, "sampled rows carry no gradient: the check would be vacuous"
    if out.d_hidden is not None:
        got = out.d_hidden[dev_pick.to(out.d_hidden.device)].float().cpu()
        res["d_hidden_rel"] = float((got - hs.grad).norm() / hs.grad.norm())
        assert res["d_hidden_rel"] < grad_rel, res

    # ---- sampled d_weight rows against the fp32 definition (half label ids, half arbitrary vocabulary entries)
    if out.d_weight is not None:
        g = torch.Generator().manual_seed(seed + 1)
        V = out.d_weight.shape[0]
        fl = lambda x: (x.reshape(-1) if keep is None else x[keep]).contiguous()
        lab_flat = labels.reshape(-1)
        some_labels = lab_flat[torch.randint(0, lab_flat.numel(), (8,), generator=g).to(lab_flat.device)].cpu()
        rows = torch.unique(torch.cat([some_labels, torch.randint(0, V, (8,), generator=g)]))
        # (the step ran without an entropy mask, so out.entropy is the raw per-token entropy the formula needs)
        ref = d_weight_rows_reference(hidden, weight, lab_flat, out.log_probs.reshape(-1), out.entropy.reshape(-1),
                                      fl(g_lp), fl(g_ent), T, rows)
        got = out.d_weight[rows.to(out.d_weight.device)].double()
        res["d_weight_rows_rel"] = float((got - ref.to(got.device)).norm() / ref.norm().clamp(min=1e-30))
        assert res["d_weight_rows_rel"] < grad_rel, res

    # ---- identities over the full result
    if out.d_weight is not None and out.d_hidden is not None:
        dw = out.d_weight
        V, N = dw.shape[0], hidden.shape[0]
        col = dw.sum(0, dtype=torch.float64)
        scale = dw.abs().sum(0, dtype=torch.float64).clamp(min=1e-30)
        res["colsum"] = float((col.abs() / scale).max())
        assert res["colsum"] < 3e-4, res
        lhs = sum(float((dw[v0:v0 + 8192].double() * weight[v0:v0 + 8192].to(dw.device).double()).sum())
                  for v0 in range(0, V, 8192))
        rhs, rhs_abs = 0.0, 0.0
        for r0 in range(0, N, 32768):
            prod = out.d_hidden[r0:r0 + 32768].double() * hidden[r0:r0 + 32768].to(out.d_hidden.device).double()
            rhs += float(prod.sum())
            rhs_abs += float(prod.abs().sum())
        res["inner_err"] = abs(lhs - rhs) / max(rhs_abs, 1e-30)
        assert res["inner_err"] < 1e-5, (res, lhs, rhs, rhs_abs)
    return res
