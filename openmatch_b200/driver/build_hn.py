"""``python -m openmatch.driver.build_hn`` — the hard-negative step of the reference's workflow
(``scripts/msmarco/build_hn.py`` as documented in ``docs/dr-msmarco-passage.md:139-157``) on pre-tokenised stores:

  --hn_file    TREC run of the TRAIN queries (``driver.retrieve --trec_save_path``)
  --qrels      tab separated ``qid 0 docid 1``
  --queries    ``.npy`` (+ ``.ids.txt``) of int32 token ids, see ``PretokenizedDataset``
  --collection ``.npy`` (+ ``.ids.txt``)
  --save_to    directory receiving ``splitNN.hn.jsonl`` (the format ``train_dr --train_dir`` reads)
"""
import random
from argparse import ArgumentParser

from ..mining import load_ranking, read_qrel, write_hn_shards


def main():
    parser = ArgumentParser()
    parser.add_argument("--hn_file", required=True)
    parser.add_argument("--qrels", required=True)
    parser.add_argument("--queries", required=True)
    parser.add_argument("--collection", required=True)
    parser.add_argument("--save_to", required=True)
    parser.add_argument("--truncate", type=int, default=128)
    parser.add_argument("--n_sample", type=int, default=30)
    parser.add_argument("--depth", type=int, default=200)
    parser.add_argument("--shard_size", type=int, default=45000)
    parser.add_argument("--seed", type=int, default=None, help="default: unseeded like the reference (datetime seed)")
    args = parser.parse_args()
    rng = random.Random(args.seed) if args.seed is not None else random
    qrel = read_qrel(args.qrels)
    paths = write_hn_shards(load_ranking(args.hn_file, qrel, args.n_sample, args.depth, rng), args.queries, args.collection,
                            args.save_to, args.shard_size, args.truncate)
    print("wrote %d shard(s) under %s" % (len(paths), args.save_to))


if __name__ == "__main__":
    main()
