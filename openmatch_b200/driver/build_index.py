"""``python -m openmatch.driver.build_index``: encode the corpus (reference: ``src/openmatch/driver/build_index.py``).
Writes ``embeddings.corpus.rank.{r}`` under ``--output_dir`` in the reference's pickle format."""
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
    config = load_config(model_args)
    tokenizer = load_tokenizer(model_args)
    model = DRModelForInference.build(model_args=model_args, config=config, cache_dir=model_args.cache_dir)
    corpus_dataset = InferenceDataset.load(tokenizer=tokenizer, data_args=data_args, is_query=False, stream=True,
                                           batch_size=encoding_args.per_device_eval_batch_size,
                                           num_processes=encoding_args.world_size,
                                           process_index=encoding_args.process_index, cache_dir=model_args.cache_dir)
    Retriever.build_embeddings(model, corpus_dataset, encoding_args)


if __name__ == '__main__':
    main()
