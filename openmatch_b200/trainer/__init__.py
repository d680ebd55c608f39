from .dense_trainer import DRTrainer, GCDenseTrainer

__all__ = ["DRTrainer", "GCDenseTrainer"]
