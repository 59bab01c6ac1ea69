"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's policy-update arithmetic.

Nothing under ``skyrl_b200/`` imports this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only as the checker / the
CPU baseline that is timed next to the CUDA path.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference from ``/root/reference``
(with stub ``ray`` / ``omegaconf`` / ``skyrl_train.config`` modules) and stores its outputs on seeded inputs
in ``tests/golden/*.npz``; ``tests/test_oracle_cpu.py`` checks this restatement against those files and against
the known-answer vectors of the reference's own unit tests.  The one unpinned piece is the numerics of
``logprobs_from_logits`` on bf16 logits: the reference delegates to flash-attn's Triton kernel and mocks it
out in its tests, so the contract there is the mathematical definition (fp32 log-softmax + gather).
"""
