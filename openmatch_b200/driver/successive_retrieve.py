"""``python -m openmatch.driver.successive_retrieve``: like ``driver.retrieve`` but searches the corpus one
embedding partition (``embeddings.corpus.rank.*``) at a time and merges the per-partition rankings, for corpora
larger than the index memory (reference: ``src/openmatch/driver/successive_retrieve.py``, which differs from its
``retrieve.py`` only in the retriever class and in not sharding the query set across processes)."""
import logging

from ..arguments import DataArguments, InferenceArguments as EncodingArguments, ModelArguments
from ..dataset import InferenceDataset
from ..modeling import DRModelForInference
from ..retriever import SuccessiveRetriever
from ..utils import save_as_trec
from ._common import load_config, load_tokenizer, parse, setup_logging

logger = logging.getLogger(__name__)


def main():
    model_args, data_args, encoding_args = parse((ModelArguments, DataArguments, EncodingArguments))
    setup_logging(encoding_args, logger)
    logger.info("Encoding parameters %s", encoding_args)
    logger.info("MODEL parameters %s", model_args)
    config = load_config(model_args)
    tokenizer = load_tokenizer(model_args, use_fast=False)
    model = DRModelForInference.build(model_args=model_args, config=config, cache_dir=model_args.cache_dir)
    query_dataset = InferenceDataset.load(tokenizer=tokenizer, data_args=data_args, is_query=True, stream=True,
                                          cache_dir=model_args.cache_dir)
    retriever = SuccessiveRetriever.from_embeddings(model, encoding_args)
    result = retriever.retrieve(query_dataset, topk=encoding_args.retrieve_depth)
    if encoding_args.process_index == 0:
        save_as_trec(result, encoding_args.trec_save_path)


if __name__ == '__main__':
    main()
