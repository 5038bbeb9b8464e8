"""TEST INFRASTRUCTURE — CPU restatement of the reference's arithmetic for the tree-OU prior + ELBO path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import anything from this package; the product (`draupnir_asr_b200/`) never does.

PINNING.  The reference's tests hold no golden vectors, known-answer tests or tolerances for this path
(SURVEY.md §4, §8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD CONTAINER:
`tests/reference_harness.py` executes the reference's own `models.py`, `models_utils.py`, `guides.py`, `train.py`
and `utils.py` in place from /root/reference on the CPU (fp64) and `tests/golden/make_reference_golden.py` commits
what they produce (`tests/golden/ref_*.npz`: -ELBO, per-site log-probs, gradients, three training epochs,
conditional mean / draw, argmax sequences, the BLOSUM input tensors).  The oracle reproduces those vectors to
rounding (`tests/test_reference_pin.py`; bit-identical losses with one thread).  ONE layer of that run is not the
reference's: `pyro-ppl` is not installed in this image, so `import pyro` resolves to
`draupnir_asr_b200.pyro_compat` — Pyro's bookkeeping (messenger stack, the ELBO sum over sites, `AutoDelta`,
`ClippedAdam`) is therefore pinned only by the restatement in SURVEY §8c and by `tests/test_shim.py`; every
floating-point operation of the model, guide, networks and sampling is the reference's code on torch.  The
patristic builder (`ete3` / R `ape`, neither available) stays pinned by three independent algorithms only
(`tests/test_oracle.py`): for that row parity is UNPINNED by the reference.
"""
