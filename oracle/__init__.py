"""CPU oracle for the word-region attention hot path of zhengfei0908/SBA-GAN.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sba_gan_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may, and there only as the checker or the timed CPU baseline.

Parity pin: the reference has no tests and no golden vectors of its own (SURVEY.md §4,
§8c), so the oracle is pinned to *the reference itself executed in the build container*
(torch 2.11.0 CPU, fp32 and fp64): ``oracle/make_golden.py`` imports
``/root/reference/AttnGAN2/code/{GlobalAttention.py,miscc/losses.py}`` unmodified, runs
them on seeded inputs and commits inputs+outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks every oracle function against those fixtures.
"""
from .attention import (  # noqa: F401
    project_words, attn_forward, attn_backward, func_attention,
    words_similarity, words_loss, words_loss_backward, ce_tail, sent_scores, sent_loss, adain_cat,
)
from .synth import (  # noqa: F401
    synth_attention_inputs, synth_words_loss_inputs, normalised_max_err,
)
