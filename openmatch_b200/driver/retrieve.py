"""``python -m openmatch.driver.retrieve``: encode queries, search the HBM index built from the corpus
pickles, write a TREC run (reference: ``src/openmatch/driver/retrieve.py``).  ``--retrieve_depth`` exposes
the top-k the reference hard-codes to 100."""
import logging

from ..arguments import DataArguments, InferenceArguments as EncodingArguments, ModelArguments
from ..dataset import InferenceDataset
from ..modeling import DRModelForInference
from ..retriever import Retriever
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
                                          batch_size=encoding_args.per_device_eval_batch_size,
                                          num_processes=encoding_args.world_size,
                                          process_index=encoding_args.process_index, cache_dir=model_args.cache_dir)
    retriever = Retriever.from_embeddings(model, encoding_args)
    # ranked arrays -> TREC lines directly (same bytes as save_as_trec on the reference's dict, minus the
    # 7 M-entry dict-of-dicts of a top-1000 MS MARCO run)
    result = retriever.retrieve(query_dataset, topk=encoding_args.retrieve_depth, as_arrays=True)
    if encoding_args.process_index == 0:
        result.save_trec(encoding_args.trec_save_path)


if __name__ == '__main__':
    main()
