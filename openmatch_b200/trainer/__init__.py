from .dense_trainer import DRTrainer, GCDenseTrainer, get_dense_rep, split_dense_inputs

__all__ = ["DRTrainer", "GCDenseTrainer", "split_dense_inputs", "get_dense_rep"]
