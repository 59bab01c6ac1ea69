"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

    python -m oracle.make_golden

Only runs in the build container (needs /root/reference).  The fixtures are small (<1 MB total) and committed;
the CPU tests check oracle/policy_oracle.py against them and the GPU tests check the CUDA kernels against them.
Where the reference's arithmetic lives in a class method we cannot import without Ray/Hydra
(RayPPOTrainer.apply_reward_kl_penalty, PolicyWorkerBase._forward_backward_micro) the script composes the same
reference primitives in the same order and cites the lines.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cfg(**kw):
    base = dict(
        policy_loss_type="regular", loss_reduction="token_mean", eps_clip_low=0.2, eps_clip_high=0.2,
        clip_ratio_c=3.0, max_seq_len=64, use_kl_loss=True, kl_estimator_type="k3", kl_loss_coef=1e-3,
        use_entropy_loss=False, entropy_loss_coef=0.01,
        off_policy_correction=SimpleNamespace(tis_ratio_type=None, sequence_mask_metric=None),
    )
    base.update(kw)
    return SimpleNamespace(**base)


def synth_batch(g, B, A, min_len=3):
    lens = torch.randint(min_len, A + 1, (B,), generator=g)
    lens[0] = A
    mask = (torch.arange(A)[None, :] < lens[:, None])
    return lens, mask


def gen_advantages(pu):
    g = torch.Generator().manual_seed(1234)
    B, A = 16, 48
    lens, mask = synth_batch(g, B, A)
    uids = np.array(["p0"] * 4 + ["p1"] * 4 + ["p2"] * 4 + ["p3"] * 3 + ["solo"])
    perm = torch.randperm(B, generator=g).numpy()
    uids = uids[perm]
    rewards = torch.zeros(B, A)
    scal = (torch.rand(B, generator=g) < 0.4).float()
    scal[torch.from_numpy(uids == "p2")] = 1.0  # a zero-variance group
    for i in range(B):
        rewards[i, lens[i] - 1] = scal[i]
    dense = rewards - 0.01 * torch.rand(B, A, generator=g) * mask.float()  # KL-in-reward style dense rewards
    values = torch.randn(B, A, generator=g)  # including padding positions (GAE does not mask them)
    out = dict(rewards=rewards.numpy(), dense_rewards=dense.numpy(), values=values.numpy(),
               mask_i64=mask.long().numpy(), uids=uids.astype("U8"))
    fm, im = mask.float(), mask.long()
    out["grpo_std"] = pu.compute_grpo_outcome_advantage(rewards.clone(), im, uids)[0].numpy()
    out["grpo_nostd"] = pu.compute_grpo_outcome_advantage(rewards.clone(), fm, uids, grpo_norm_by_std=False)[0].numpy()
    out["grpo_dense"] = pu.compute_grpo_outcome_advantage(dense.clone(), fm, uids)[0].numpy()
    out["rloo"] = pu.compute_rloo_outcome_advantage(rewards.clone(), fm, uids)[0].numpy()
    out["rloo_dense"] = pu.compute_rloo_outcome_advantage(dense.clone(), im, uids)[0].numpy()
    for name, (gm, lm) in {"gae_1_095": (1.0, 0.95), "gae_099_1": (0.99, 1.0), "gae_1_1": (1.0, 1.0)}.items():
        adv, ret = pu.compute_gae_advantage_return(dense, values, fm, gamma=gm, lambd=lm)
        out[name + "_adv"], out[name + "_ret"] = adv.numpy(), ret.numpy()
    for gm in (1.0, 0.9):
        adv, ret = pu.compute_reinforce_plus_plus_outcome_advantage(dense, fm, gamma=gm)
        out[f"rpp_{gm}_adv"], out[f"rpp_{gm}_ret"] = adv.numpy(), ret.numpy()
    hole = fm.clone()
    hole[2, 3:6] = 0.0  # a hole inside the response: the running return resets there
    adv, ret = pu.compute_reinforce_plus_plus_outcome_advantage(dense, hole, gamma=0.95)
    out["rpp_hole_mask"], out["rpp_hole_adv"], out["rpp_hole_ret"] = hole.numpy(), adv.numpy(), ret.numpy()
    batch = {"advantages": torch.from_numpy(out["gae_1_095_adv"]).clone() * 3 + 0.5, "response_mask": fm}
    out["norm_in"] = batch["advantages"].numpy().copy()
    out["norm_out"] = pu.normalize_advantages_dict(batch)["advantages"].numpy()
    # KL in reward: trainer.py:992-1011 composed from the reference primitives
    lp = -torch.rand(B, A, generator=g) * 2
    base = lp + 0.1 * torch.randn(B, A, generator=g)
    base[1, 2] = lp[1, 2] + 25.0  # clamp hits
    base[2, 3] = lp[2, 3] - 25.0
    out["lp"], out["base"] = lp.numpy(), base.numpy()
    for kind in ("k1", "abs", "k2", "k3"):
        kl = pu.compute_approx_kl(lp, base, loss_mask=fm, kl_estimator_type=kind)
        out["kl_" + kind] = kl.numpy()
        out["kl_nomask_" + kind] = pu.compute_approx_kl(lp, base, loss_mask=None, kl_estimator_type=kind).numpy()
        kl_max = torch.max(kl.abs(), dim=-1)[0]
        kl_mean = pu.masked_mean(kl, fm, dim=-1)
        out["klrew_" + kind] = (rewards - kl * max(0, 0.05)).numpy()
        out["klrew_metrics_" + kind] = np.array([kl_mean.mean().item(), kl_max.mean().item()], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "advantages.npz"), **out)


def gen_losses(pu):
    g = torch.Generator().manual_seed(1235)
    B, A = 12, 40
    lens, mask = synth_batch(g, B, A)
    fm = mask.float()
    fm[3] = 0.0  # a fully masked row
    fm[5, 4:9] = 0.0  # an interior hole (multi-turn tool tokens)
    lp = -torch.rand(B, A, generator=g) * 3
    olp = lp + 0.05 * torch.randn(B, A, generator=g)
    olp[0, 0] = lp[0, 0] + 25.0  # ratio clamp low
    olp[0, 1] = lp[0, 1] - 25.0  # ratio clamp high
    olp[1, :6] = lp[1, :6]  # exact ties: ratio == 1
    olp[2, 5] = lp[2, 5] - 0.5  # clipped high
    olp[2, 6] = lp[2, 6] + 0.5  # clipped low
    adv = torch.randn(B, A, generator=g)
    adv[1, 2] = 0.0
    adv[4] = -adv[4].abs() * 3  # strongly negative advantages for dual clip
    base = lp + 0.1 * torch.randn(B, A, generator=g)
    ent = torch.rand(B, A, generator=g) * 2
    out = dict(lp=lp.numpy(), olp=olp.numpy(), adv=adv.numpy(), mask=fm.numpy(), base=base.numpy(), ent=ent.numpy())
    for kind in ("regular", "dual_clip", "gspo"):
        for red in ("token_mean", "sequence_mean", "seq_mean_token_sum_norm"):
            for mname, m in (("mask", fm), ("nomask", None)):
                c = cfg(policy_loss_type=kind, loss_reduction=red, eps_clip_high=0.28 if kind == "dual_clip" else 0.2)
                x = lp.clone().requires_grad_(True)
                loss, met = pu.PolicyLossRegistry.get(kind)(x, olp, adv, c, loss_mask=m)
                loss.backward()
                key = f"{kind}__{red}__{mname}"
                out[key + "__loss"] = np.float32(loss.item())
                out[key + "__clip"] = np.float32(met["clip_ratio"])
                out[key + "__grad"] = x.grad.numpy()
    # the rest of the elementwise loss registry (sapo / cispo / cross_entropy / importance_sampling)
    xcfg = dict(sapo=SimpleNamespace(tau_pos=1.0, tau_neg=2.0),
                cispo=SimpleNamespace(cispo_eps_clip_low=0.1, cispo_eps_clip_high=0.3))
    olp_mild = lp + 0.2 * torch.randn(B, A, generator=g)  # importance_sampling has no clamp: keep ratios finite
    out["olp_mild"] = olp_mild.numpy()
    for kind, red in (("sapo", "sequence_mean"), ("sapo", "token_mean"), ("cispo", "token_mean"),
                      ("cispo", "seq_mean_token_sum_norm"), ("cross_entropy", "token_mean"),
                      ("importance_sampling", "token_mean")):
        for mname, m in (("mask", fm), ("nomask", None)):
            c = cfg(policy_loss_type=kind, loss_reduction=red, **xcfg)
            x = lp.clone().requires_grad_(True)
            o = olp_mild if kind == "importance_sampling" else olp
            loss, met = pu.PolicyLossRegistry._functions[kind](x, o, adv, c, loss_mask=m)
            loss.backward()
            key = f"{kind}__{red}__{mname}"
            out[key + "__loss"] = np.float32(loss.item())
            out[key + "__metric"] = np.float32(list(met.values())[0])
            out[key + "__grad"] = x.grad.numpy()
    # covariance-selection losses (ppo_utils.py:765-888).  clip_cov draws torch.randperm from the global CPU generator:
    # seed it here, the oracle test seeds it the same way and replays the draw.
    cov_cfg = dict(clip_cov=SimpleNamespace(clip_ratio=0.1, clip_cov_lb=0.05, clip_cov_ub=5.0),
                   kl_cov=SimpleNamespace(kl_cov_frac=0.2, ppo_kl_coef=1.0))
    for kind, red in (("clip_cov", "token_mean"), ("clip_cov", "sequence_mean"), ("kl_cov", "token_mean"),
                      ("kl_cov", "seq_mean_token_sum_norm")):
        c = cfg(policy_loss_type=kind, loss_reduction=red, **cov_cfg)
        x = lp.clone().requires_grad_(True)
        torch.manual_seed(777)
        loss, met = pu.PolicyLossRegistry._functions[kind](x, olp_mild, adv, c, loss_mask=fm)
        loss.backward()
        key = f"{kind}__{red}__mask"
        out[key + "__loss"] = np.float32(loss.item())
        out[key + "__metric"] = np.float32(met["clip_ratio"])
        out[key + "__grad"] = x.grad.numpy()
    # off-policy correction inside ppo_policy_loss (off_policy_correction_utils.py)
    rlp = olp + 0.15 * torch.randn(B, A, generator=g)
    rlp[6, 3] = olp[6, 3] - 3.0  # one outlier token (ratio e^3) -> its whole sequence is masked by the outlier rule
    rlp[7] = olp[7] - 0.05       # a sequence with a uniformly higher ratio
    out["rlp"] = rlp.numpy()
    opc_cases = {
        "token_tis": dict(tis_ratio_type="token", token_tis_ratio_clip_high=1.2),
        "seq_tis_outlier": dict(tis_ratio_type="sequence", sequence_tis_ratio_clip_high=2.0,
                                outlier_token_is_threshold_low=0.2, outlier_token_is_threshold_high=5.0),
        "geo_mask": dict(sequence_mask_metric="geometric", geo_mask_low=0.98, geo_mask_high=1.02),
        "product_mask_token_tis": dict(tis_ratio_type="token", sequence_mask_metric="product", product_mask_low=0.5,
                                       product_mask_high=2.0, outlier_token_is_threshold_high=10.0),
    }
    defaults = dict(tis_ratio_type=None, token_tis_ratio_clip_high=2.0, sequence_tis_ratio_clip_high=5.0,
                    sequence_mask_metric=None, geo_mask_high=1.01, geo_mask_low=0.99, product_mask_high=2.0,
                    product_mask_low=0.5, outlier_token_is_threshold_low=None, outlier_token_is_threshold_high=None)
    import json as _json
    for name, kw in opc_cases.items():
        okw = dict(defaults); okw.update(kw)
        for kind, red in (("regular", "token_mean"), ("dual_clip", "sequence_mean")):
            c = cfg(policy_loss_type=kind, loss_reduction=red, off_policy_correction=SimpleNamespace(**okw))
            x = lp.clone().requires_grad_(True)
            loss, met = pu.PolicyLossRegistry.get(kind)(x, olp, adv, c, loss_mask=fm, rollout_logprobs=rlp)
            loss.backward()
            key = f"opc__{name}__{kind}"
            out[key + "__loss"] = np.float32(loss.item())
            out[key + "__grad"] = x.grad.numpy()
            out[key + "__metrics"] = np.array(_json.dumps(met))
            out[key + "__cfg"] = np.array(_json.dumps(okw))
    # critic loss (ppo_utils.py:175-193)
    vals = torch.randn(B, A, generator=g)
    old_vals = vals + 0.3 * torch.randn(B, A, generator=g)
    old_vals[0, :4] = vals[0, :4]  # exact ties surr1 == surr2
    rets = torch.randn(B, A, generator=g)
    out.update(values=vals.numpy(), old_values=old_vals.numpy(), returns=rets.numpy())
    for cname, clip in (("clip02", 0.2), ("noclip", None)):
        for mname, m in (("mask", fm), ("nomask", None)):
            v = vals.clone().requires_grad_(True)
            loss, clipfrac = pu.ppo_critic_loss(v, old_vals, rets, SimpleNamespace(value_clip=clip), loss_mask=m)
            loss.backward()
            key = f"critic__{cname}__{mname}"
            out[key + "__loss"] = np.float32(loss.item())
            out[key + "__clipfrac"] = np.float32(-1.0 if clipfrac is None else clipfrac)
            out[key + "__grad"] = v.grad.numpy()
    # worker.py:787-889 loss assembly from the reference primitives, same order of operations
    for name, kw in {
        "asm_grpo_kl": dict(policy_loss_type="regular", use_kl_loss=True, use_entropy_loss=False),
        "asm_ent": dict(policy_loss_type="dual_clip", use_kl_loss=False, use_entropy_loss=True, clip_ratio_c=3.0),
        "asm_gspo_all": dict(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=True,
                             use_entropy_loss=True, kl_estimator_type="k2"),
    }.items():
        c = cfg(**kw)
        x = lp.clone().requires_grad_(True)
        e = ent.clone().requires_grad_(True)
        policy_loss, met = pu.PolicyLossRegistry.get(c.policy_loss_type)(x, olp, adv, config=c, loss_mask=fm)
        with torch.set_grad_enabled(c.use_entropy_loss):
            entropy = pu.masked_mean(e, fm)
        ent_term = entropy * c.entropy_loss_coef if c.use_entropy_loss else torch.tensor(0.0)
        if c.use_kl_loss:
            kl = pu.compute_approx_kl(x, base, loss_mask=fm, kl_estimator_type=c.kl_estimator_type)
            kl = pu.masked_mean(kl, fm, dim=-1).mean()
        else:
            kl = torch.tensor(0.0)
        final = policy_loss + kl * c.kl_loss_coef - ent_term
        final.backward()
        out[name + "__stats"] = np.array([final.item(), policy_loss.item(), entropy.item(), kl.item(),
                                          met["clip_ratio"]], dtype=np.float32)
        out[name + "__dlp"] = x.grad.numpy()
        out[name + "__dent"] = (e.grad if e.grad is not None else torch.zeros_like(e)).numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def gen_head(pu, tu):
    g = torch.Generator().manual_seed(1236)
    N, H, V = 80, 64, 640  # V is not a multiple of the 256-wide vocab slab
    hidden = torch.randn(N, H, generator=g).to(torch.bfloat16)
    weight = (torch.randn(V, H, generator=g) * (2.0 / H ** 0.5)).to(torch.bfloat16)
    labels = torch.randint(0, V, (N,), generator=g)
    out = dict(hidden=hidden.float().numpy(), weight=weight.float().numpy(), labels=labels.numpy())
    for T in (1.0, 0.6):
        h = hidden.float().requires_grad_(True)
        w = weight.float().requires_grad_(True)
        logits = F.linear(h, w) / T  # model_wrapper.py:309-314 in fp32
        lp = tu.logprobs_from_logits_v2(logits, labels)  # torch_utils.py:158-177 (fp32 branch)
        ent = tu.chunked_entropy_from_logits(logits.unsqueeze(0), requires_grad=True).squeeze(0)
        g_lp = torch.randn(N, generator=g)
        g_ent = torch.randn(N, generator=g) * 0.1
        (lp * g_lp).sum().backward(retain_graph=True)
        dh_lp, dw_lp = h.grad.clone(), w.grad.clone()
        h.grad = None
        w.grad = None
        ((lp * g_lp).sum() + (ent * g_ent).sum()).backward()
        k = f"T{T}"
        out.update({k + "_lp": lp.detach().numpy(), k + "_ent": ent.detach().numpy(), k + "_g_lp": g_lp.numpy(),
                    k + "_g_ent": g_ent.numpy()})
        if T == 1.0:  # log-prob gradient only
            out.update({k + "_dh_lp": dh_lp.numpy(), k + "_dw_lp": dw_lp.numpy()})
        else:  # log-prob + entropy gradient
            out.update({k + "_dh_both": h.grad.numpy(), k + "_dw_both": w.grad.numpy()})
    # logits helpers on given logits (torch_utils.py), with an attention mask
    V = 1000
    logits = torch.randn(2, 12, V, generator=g) * 2
    labels2 = torch.randint(0, V, (2, 12), generator=g)
    amask = (torch.rand(2, 12, generator=g) > 0.3).float()
    out["logits"] = logits.numpy()
    out["labels2"] = labels2.numpy()
    out["amask"] = amask.numpy()
    out["logits_lp"] = tu.logprobs_from_logits_v2(logits, labels2).numpy()
    out["logits_ent"] = tu.chunked_entropy_from_logits(logits, attention_mask=amask).numpy()
    z = logits.clone().requires_grad_(True)
    u = torch.randn(2, 12, generator=g)
    v = torch.randn(2, 12, generator=g)
    (tu.logprobs_from_logits_v2(z, labels2) * u).sum().backward()
    out["u"], out["v"] = u.numpy(), v.numpy()
    out["dlogits_lp"] = z.grad.numpy()
    z2 = logits.clone().requires_grad_(True)
    (tu.chunked_entropy_from_logits(z2, requires_grad=True, attention_mask=amask) * v).sum().backward()
    out["dlogits_ent"] = z2.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "head.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    pu, tu = ref_loader.load()
    torch.set_num_threads(4)
    gen_advantages(pu)
    gen_losses(pu)
    gen_head(pu, tu)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
