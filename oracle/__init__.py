"""TEST INFRASTRUCTURE — CPU restatement of the reference's arithmetic for the tree-OU prior + ELBO path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import anything from this package; the product (`draupnir_asr_b200/`) never does.

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or tolerances for this path
 and cannot be imported here (pyro, ignite, Bio, ete3, matplotlib and R are absent), so
this restatement — the same `torch.distributions` / `torch.linalg` / `nn.GRU` calls the reference makes,
in fp64 — *is* the oracle.  It is cross-checked against closed forms, against an independent Schur-complement
derivation and against brute-force tree walks in `tests/test_oracle.py`.
"""
