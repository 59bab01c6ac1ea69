"""Known-answer vectors of the reference's off-policy-correction tests
(skyrl-train/tests/cpu/algorithms/test_off_policy_correction.py:20-330), shared by the CPU and GPU suites.
Each case: (old_log_probs, rollout_logprobs, loss_mask, config overrides, expected tis or None, expected new mask
row multiplier, expected metrics subset)."""
from types import SimpleNamespace

DEFAULTS = dict(tis_ratio_type=None, token_tis_ratio_clip_high=2.0, sequence_tis_ratio_clip_high=5.0,
                sequence_mask_metric=None, geo_mask_high=1.01, geo_mask_low=0.99, product_mask_high=2.0,
                product_mask_low=0.5, outlier_token_is_threshold_low=None, outlier_token_is_threshold_high=None)


def cfg(**kw):
    d = dict(DEFAULTS)
    d.update(kw)
    return SimpleNamespace(**d)


CASES = [
    # name, olp, rlp, mask, cfg, expected tis (per token) or None, expected row keep, metrics
    ("token_tis_cap", [[-1.0, -1.5, -0.5]], [[-1.5, -1.0, -1.5]], [[1.0, 1.0, 1.0]],
     dict(tis_ratio_type="token", token_tis_ratio_clip_high=2.0), [[1.6487, 0.6065, 2.0]], [1.0],
     {"tis_token_clip_high_ratio": 1 / 3}),
    ("seq_tis", [[-1.0, -1.5, -0.5]], [[-1.5, -1.0, -1.5]], [[1.0, 1.0, 1.0]],
     dict(tis_ratio_type="sequence", sequence_tis_ratio_clip_high=5.0), [[2.7183] * 3], [1.0],
     {"tis_seq_clip_high_ratio": 0.0}),
    ("seq_tis_capped", [[-1.0, -1.0, -1.0]], [[-2.0, -2.0, -2.0]], [[1.0, 1.0, 1.0]],
     dict(tis_ratio_type="sequence", sequence_tis_ratio_clip_high=5.0), [[5.0] * 3], [1.0],
     {"tis_seq_clip_high_ratio": 1.0}),
    ("seq_tis_masked_token", [[-1.0, -1.5, -0.5]], [[-1.5, -1.0, -1.5]], [[1.0, 0.0, 1.0]],
     dict(tis_ratio_type="sequence", sequence_tis_ratio_clip_high=10.0), [[4.4817] * 3], [1.0],
     {"tis_seq_clip_high_ratio": 0.0}),
    ("geo_accept", [[-1.0, -1.1, -1.0]], [[-1.1, -1.0, -1.0]], [[1.0, 1.0, 1.0]],
     dict(sequence_mask_metric="geometric", geo_mask_high=1.1, geo_mask_low=0.9), None, [1.0],
     {"geo_sequence_mask_masked_ratio": 0.0, "geo_sequence_mask_over_high_ratio": 0.0,
      "geo_sequence_mask_under_low_ratio": 0.0}),
    ("geo_reject", [[-1.0, -1.0, -1.0]], [[-1.5, -1.5, -1.5]], [[1.0, 1.0, 1.0]],
     dict(sequence_mask_metric="geometric", geo_mask_high=1.1, geo_mask_low=0.9), None, [0.0],
     {"geo_sequence_mask_masked_ratio": 1.0, "geo_sequence_mask_over_high_ratio": 1.0,
      "geo_sequence_mask_under_low_ratio": 0.0}),
    ("product_accept", [[-1.0, -1.1, -1.2]], [[-1.2, -1.2, -1.2]], [[1.0, 1.0, 1.0]],
     dict(sequence_mask_metric="product", product_mask_high=2.0, product_mask_low=0.5), None, [1.0],
     {"product_sequence_mask_masked_ratio": 0.0}),
    ("product_reject", [[-1.0, -1.0, -1.0]], [[-2.0, -2.0, -2.0]], [[1.0, 1.0, 1.0]],
     dict(sequence_mask_metric="product", product_mask_high=2.0, product_mask_low=0.5), None, [0.0],
     {"product_sequence_mask_masked_ratio": 1.0, "product_sequence_mask_over_high_ratio": 1.0}),
    ("outlier_reject", [[-1.0, -1.0, -1.0]], [[-1.0, -1.0, -6.0]], [[1.0, 1.0, 1.0]],
     dict(tis_ratio_type="token", token_tis_ratio_clip_high=1e9, outlier_token_is_threshold_low=1e-4,
          outlier_token_is_threshold_high=100.0), None, [0.0],
     {"outlier_seq_masked_ratio": 1.0, "outlier_seq_over_high_ratio": 1.0, "outlier_seq_under_low_ratio": 0.0}),
    ("outlier_masked_token_ignored", [[-1.0, -1.0, -1.0]], [[-1.0, -1.0, -6.0]], [[1.0, 1.0, 0.0]],
     dict(tis_ratio_type="token", token_tis_ratio_clip_high=1e9, outlier_token_is_threshold_low=1e-4,
          outlier_token_is_threshold_high=100.0), None, [1.0], {"outlier_seq_masked_ratio": 0.0}),
    ("outlier_null_thresholds", [[-1.0, -1.0, -1.0]], [[-1.0, -1.0, -6.0]], [[1.0, 1.0, 1.0]],
     dict(tis_ratio_type="token", token_tis_ratio_clip_high=1e9), None, [1.0], {"outlier_seq_masked_ratio": 0.0}),
]
