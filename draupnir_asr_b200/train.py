"""`train` with the reference's signature (S/train.py:58-79): one `svi.step` over the whole tree per epoch."""
from __future__ import annotations


def train(svi, training_function_input):
    patristic_matrix = training_function_input["patristic_matrix_model"]
    cladistic_matrix = training_function_input["cladistic_matrix_full"]
    dataset_blosum = training_function_input["dataset_train_blosum"]
    train_loader = training_function_input["train_loader"]
    map_estimates = training_function_input["map_estimates"]
    train_loss = 0.0
    for batch_number, dataset in enumerate(train_loader):
        train_loss += svi.step(dataset, patristic_matrix, cladistic_matrix, dataset_blosum, None, map_estimates)
    return train_loss


def select_training_function(batch_by_clade: bool = False, batch_size=1):
    """S/train.py:117-145: only the whole-tree path is accelerated."""
    if batch_by_clade or (batch_size is not None and batch_size > 1):
        raise NotImplementedError("batched / clade-batched training approximates the OU process by independent blocks "
                                  "and is outside the accelerated path")
    return train
