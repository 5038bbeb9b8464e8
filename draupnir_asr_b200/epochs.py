"""The epoch loop around `svi.step`, slimmed.

The reference's loop (S/main.py:1089-1128; S/ = /root/reference/draupnir/src/draupnir/) does, EVERY epoch: a second guide
forward to obtain `map_estimates` (:1097), one `svi.step` (:1100), an argmax decode of all leaves (:1113-1120: one more
decoder forward), two checkpoint writes (:1121-1122), and the per-site entropies / percent identity of the decoded leaves
(:1124-1128, the latter on the CPU).  Measured on this implementation (`profiles/epoch_extras.py`,
`profiles/r01/epoch_extras.txt`) these extras cost 0.4-0.55 of a step before any disk I/O (C4: 6.2 ms second guide call +
3.1 ms decode + 1.8 ms serialisation against a 23.3 ms step), so the loop below keeps their RESULTS and drops their repetition:

* `map_estimates` are the guide's return value of the step that was just taken (`Trace_ELBO.last_traces`: no second guide
  forward; the entries that need the encoder's complete pass stay deferred until somebody reads them);
* the argmax decode + entropies + percent identity run every `stats_every` epochs (1 = the reference's cadence), on the GPU;
* checkpoints (`Model_state_dict.p`, `Guide_state_dict.p`, `Optimizer_state.p`: S/main.py:328-342) are written every
  `check_point_epoch` epochs (1 = the reference's cadence).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch


def calculate_percent_id(dataset_true: torch.Tensor, aa_sequences_predictions: torch.Tensor, align_lenght: int):
    """(mean, std) over sequences of the percentage of sites where the prediction equals the observed residue
    (S/main.py:527-539); stays on the device until the two scalars are read."""
    equal = (aa_sequences_predictions == dataset_true[:, 2:, 0]).double()
    pid = equal.sum(-1) / align_lenght * 100
    return float(pid.mean()), float(pid.std())


def compute_sites_entropies(logits: torch.Tensor, node_names: torch.Tensor) -> torch.Tensor:
    """`[n, 1 + L]`: node id followed by the per-site entropy of `prob = exp(logit) / (1 + exp(logit))`
    (S/models_utils.py:409-427, the reference's own definition, kept as is)."""
    prob = torch.sigmoid(logits)
    ent = -torch.sum(prob * torch.log(prob), dim=2)
    return torch.cat((node_names[:, None].to(ent.dtype), ent), dim=1)


def save_checkpoint(Draupnir, save_directory: str, optimizer) -> None:
    """S/main.py:328-335."""
    d = os.path.join(save_directory, "Draupnir_Checkpoints")
    os.makedirs(d, exist_ok=True)
    optimizer.save(os.path.join(d, "Optimizer_state.p"))
    torch.save(Draupnir.state_dict(), os.path.join(d, "Model_state_dict.p"))


def save_checkpoint_guide(guide, save_directory: str) -> None:
    """S/main.py:336-341."""
    d = os.path.join(save_directory, "Draupnir_Checkpoints")
    os.makedirs(d, exist_ok=True)
    torch.save(guide.state_dict(), os.path.join(d, "Guide_state_dict.p"))


def _guide_return(svi, guide, step_args):
    """`map_estimates` of the step just taken: the guide's recorded return value, else one guide call (real Pyro's traces
    keep it under `_RETURN`; an ELBO object without `last_traces` costs the extra forward the reference pays)."""
    traces = getattr(svi.loss, "last_traces", None)
    if traces is not None:
        ret = getattr(traces[1], "return_value", None)
        if ret is None and hasattr(traces[1], "nodes") and "_RETURN" in traces[1].nodes:
            ret = traces[1].nodes["_RETURN"]["value"]
        if isinstance(ret, dict):
            return ret
    with torch.no_grad():
        return guide(*step_args)


def run_epochs(svi, Draupnir, guide, step_args, patristic_matrix_full, num_epochs: int, *, results_dir: Optional[str] = None,
               stats_every: int = 1, check_point_epoch: int = 1, on_epoch: Optional[Callable[[int, Dict], None]] = None) -> Dict:
    """`num_epochs` SVI steps over the whole tree with the reference's per-epoch bookkeeping (see the module docstring).
    `step_args` are the six positional arguments of `svi.step` (S/train.py:75).  Returns the lists the reference plots
    (`train_loss`, `average_pid`, `std_pid`), the last `entropies` tensor and the last `map_estimates`."""
    dataset_train = step_args[0]
    node_ids = dataset_train[:, 0, 1]
    align_len = dataset_train.shape[1] - 2
    out: Dict[str, List] = {"train_loss": [], "average_pid": [], "std_pid": [], "stats_epochs": []}
    entropies = map_estimates = None
    for epoch in range(num_epochs):
        loss = svi.step(*step_args)
        out["train_loss"].append(float(loss))
        map_estimates = _guide_return(svi, guide, step_args)
        if stats_every and epoch % stats_every == 0:
            # plain dict of what is already there (deferred entries of the variational guide stay untouched)
            raw = ((k, dict.__getitem__(map_estimates, k)) for k in map_estimates)
            estimates = {k: v.detach() for k, v in raw if isinstance(v, torch.Tensor)}
            sample = Draupnir.sample(estimates, 1, dataset_train, patristic_matrix_full, None, use_argmax=True, use_test=False,
                                     use_test2=False)
            entropies = compute_sites_entropies(sample.logits, node_ids)
            pid, std = calculate_percent_id(dataset_train, sample.aa_sequences, align_len)
            out["average_pid"].append(pid)
            out["std_pid"].append(std)
            out["stats_epochs"].append(epoch)
        if results_dir is not None and check_point_epoch and epoch % check_point_epoch == 0:
            save_checkpoint(Draupnir, results_dir, svi.optim)
            if hasattr(guide, "state_dict"):
                save_checkpoint_guide(guide, results_dir)
        if on_epoch is not None:
            on_epoch(epoch, {"loss": out["train_loss"][-1], "map_estimates": map_estimates})
    out["entropies"] = entropies
    out["map_estimates"] = map_estimates
    return out
