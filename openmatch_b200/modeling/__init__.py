from .dense_retrieval_model import DRModel, DRModelForInference, DROutput
from .linear import LinearHead

__all__ = ["DRModel", "DRModelForInference", "DROutput", "LinearHead"]
