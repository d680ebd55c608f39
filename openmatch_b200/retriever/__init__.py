from .dense_retriever import FaissRetriever, Retriever, SuccessiveRetriever

__all__ = ["Retriever", "SuccessiveRetriever", "FaissRetriever"]
