"""Shared CLI glue of the three hot-path drivers."""
import logging
import os
import sys

from transformers import HfArgumentParser


def parse(dataclasses_):
    parser = HfArgumentParser(dataclasses_)
    if len(sys.argv) == 2 and sys.argv[1].endswith(".json"):
        return parser.parse_json_file(json_file=os.path.abspath(sys.argv[1]))
    return parser.parse_args_into_dataclasses()


def setup_logging(run_args, logger):
    logging.basicConfig(format="%(asctime)s - %(levelname)s - %(name)s -   %(message)s", datefmt="%m/%d/%Y %H:%M:%S",
                        level=logging.INFO if run_args.local_rank in [-1, 0] else logging.WARN)
    logger.warning("Process rank: %s, device: %s, n_gpu: %s, distributed: %s, fp16: %s, bf16: %s", run_args.local_rank,
                   run_args.device, run_args.n_gpu, bool(run_args.local_rank != -1), run_args.fp16, run_args.bf16)


def load_tokenizer(model_args, **kw):
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(model_args.tokenizer_name or model_args.model_name_or_path,
                                         cache_dir=model_args.cache_dir, **kw)


def load_config(model_args):
    from transformers import AutoConfig
    path = model_args.config_name or model_args.model_name_or_path
    # An UNTIED OpenMatch checkpoint keeps its HF configs under query_model/ and passage_model/ (DRModel.save,
    # dense_retrieval_model.py:230-245) and has no config.json at the root, which the reference's drivers trip over
    # (AutoConfig.from_pretrained(root), build_index.py:24-28) unless --config_name is given: resolve it here.
    if os.path.isdir(path) and not os.path.exists(os.path.join(path, "config.json")) and \
            os.path.exists(os.path.join(path, "passage_model", "config.json")):
        path = os.path.join(path, "passage_model")
    return AutoConfig.from_pretrained(path, num_labels=1, cache_dir=model_args.cache_dir)
