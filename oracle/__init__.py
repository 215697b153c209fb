"""CPU oracle for the IBC energy-based-policy hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``ibc_b200/`` imports it, and the product
path raises when the CUDA library is missing instead of falling back to this.

It is a restatement (PyTorch-CPU for the floating-point parts, NumPy + plain C
for the fp64 multinomial/bincount integer part) of the reference algorithms:

  ibc/agents/mcmc.py                (all)            -> oracle/mcmc.py
  ibc/losses/ebm_loss.py:21-48      info_nce         -> oracle/losses.py
  ibc/losses/gradient_loss.py:23-72 grad_penalty     -> oracle/losses.py
  networks/mlp_ebm.py:72-96, networks/layers/resnet.py:135-153,
  networks/layers/spectral_norm.py:57-86             -> oracle/mlp_ebm.py
  networks/pixel_ebm.py:79-125, layers/conv_maxpool.py:20-48,
  layers/dense_resnet_value.py:22-69, utils/image_prepro.py:20-43
                                                     -> oracle/pixel_ebm.py
  ibc/agents/ibc_agent.py:132-300,340-475, ibc/agents/base_agent.py:36-62,
  ibc/train/get_agent.py:64-68 (Keras Adam + ExponentialDecay)
                                                     -> oracle/agent.py
  ibc/agents/ibc_policy.py:240-328                   -> oracle/agent.py

PARITY PINNING.  The reference is TensorFlow 2.6 / tf-agents / tfp Python and
cannot be imported in this image (no TF, no network), and it ships no golden
vectors or known-answer tests for this path.  What pins this oracle:
  * the reference's own behavioural tests (ibc/agents/mcmc_test.py:45-57,
    117-144, 146-170) re-run against the oracle in tests/test_oracle_*.py,
  * the hand-derived known answers of SURVEY.md Appendix B (InfoNCE epsilon
    form, polynomial schedule values, one Langevin step, resample indices,
    DFO probs/actions misalignment, tile_batch order, image stacking quirk,
    grad-penalty on a linear net),
  * cross-checks of the manual backward / tangent formulas against torch
    autograd (first and second order).
The arithmetic that lives in third-party code (tf.random.categorical's fp64
sequential CDF + upper_bound, Keras KLDivergence epsilon clipping, Keras Adam)
is restated from the published TF 2.6.0 / Keras 2.6.0 sources; it cannot be
re-verified offline.  Bit-level TF parity is therefore **unpinned**; the
bit-exact claim is GPU == this oracle on identical logits and uniform draws.
"""
