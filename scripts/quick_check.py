"""A few seconds of GPU time: smoke(), then the code paths that were written after the last full GPU suite - the four
remaining registry losses through the fused step (loss_flat_kernel<-1>) and GSPO with chunks off the tile grid.
Prints one line per check as it goes (a killed run still leaves a partial log in gpurun_out/quick_check.log)."""
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "quick_check.log"), "a")


def say(msg):
    line = f"[{time.time() - t0:6.1f}s] {msg}"
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


import torch  # noqa: E402

say(f"torch imported, cuda={torch.cuda.is_available()}")
import __graft_entry__ as entry  # noqa: E402

try:
    entry.smoke()
    say("smoke OK")
except Exception as exc:  # noqa: BLE001
    say(f"smoke FAILED: {exc!r}")
import test_zconfigs_gpu as z  # noqa: E402

checks = [("fused sapo", lambda: z.test_fused_step_runs_the_rest_of_the_registry("sapo", "sequence_mean")),
          ("fused cispo", lambda: z.test_fused_step_runs_the_rest_of_the_registry("cispo", "token_mean")),
          ("fused importance_sampling", lambda: z.test_fused_step_runs_the_rest_of_the_registry("importance_sampling", "token_mean")),
          ("fused cross_entropy", lambda: z.test_fused_step_runs_the_rest_of_the_registry("cross_entropy", "token_mean")),
          ("gspo off-grid A=1000", lambda: z.test_gspo_fused_step_with_a_response_length_off_the_tile_grid(1000)),
          ("gspo off-grid A=1001", lambda: z.test_gspo_fused_step_with_a_response_length_off_the_tile_grid(1001)),
          ("gspo all-dead chunk 0", lambda: z.test_gspo_chunk_whose_every_row_is_dead(0)),
          ("gspo all-dead chunk 1", lambda: z.test_gspo_chunk_whose_every_row_is_dead(1)),
          ("sharded gae", z.test_sharded_gae_equals_the_batch_gae),
          ("kl gradient k3", lambda: z.test_kl_loss_term_with_the_gradient_flag("k3")),
          ("peaked logits, autograd path", lambda: z.test_sharply_peaked_logits(False)),
          ("peaked logits, fused step", lambda: z.test_sharply_peaked_logits(True))]
for name, fn in checks:
    try:
        fn()
        say(f"{name}: OK")
    except Exception as exc:  # noqa: BLE001
        say(f"{name}: FAILED {type(exc).__name__}: {str(exc)[:300]}")
say("done")
