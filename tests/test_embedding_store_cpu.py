"""CPU: the streaming embedding-file reader / writer (openmatch_b200/embedding_store.py) against plain pickle — the
reference's format (src/openmatch/retriever/dense_retriever.py:84-86,96-101) in both directions, and
scripts/split_embeddings.py interop."""
import pickle

import numpy as np
import pytest

from openmatch_b200.embedding_store import EmbeddingFile, read_embedding_file, split_embedding_file, write_embedding_file


def _reference_dump(path, enc, ids):
    with open(path, "wb") as f:
        pickle.dump((enc, ids), f, protocol=4)  # dense_retriever.py:84-86


@pytest.mark.parametrize("n,d", [(0, 8), (1, 4), (37, 16), (3000, 64), (70000, 32)])
def test_writer_output_is_what_pickle_load_expects(tmp_path, n, d):
    rng = np.random.default_rng(n + d)
    enc = rng.standard_normal((n, d), dtype=np.float32)
    ids = ["doc-%d" % i for i in range(n)]
    if n > 2:
        ids[1] = "x" * 300  # long id (BINUNICODE) and a repeated object
        ids[2] = ids[0]
    p = tmp_path / "embeddings.corpus.rank.0"
    write_embedding_file(str(p), enc, ids, chunk_rows=1024)
    with open(p, "rb") as f:
        got_enc, got_ids = pickle.load(f)  # what the reference's init_index_and_add does (:96-101)
    assert isinstance(got_enc, np.ndarray) and got_enc.dtype == np.float32 and got_enc.shape == (n, d)
    assert got_enc.flags["C_CONTIGUOUS"] and np.array_equal(got_enc, enc)
    assert isinstance(got_ids, list) and got_ids == ids


@pytest.mark.parametrize("n,d", [(0, 8), (5, 4), (3000, 64), (70000, 32)])
def test_reader_maps_reference_files_without_loading_them(tmp_path, n, d):
    rng = np.random.default_rng(n)
    enc = rng.standard_normal((n, d), dtype=np.float32)
    ids = [str(10 * i) for i in range(n)]
    if n > 3:
        ids[3] = ids[1]  # memoised duplicate object inside the list
    p = tmp_path / "ref.pkl"
    _reference_dump(p, enc, ids)
    ef = EmbeddingFile(str(p))
    assert ef.streaming and ef.shape == (n, d)
    assert np.array_equal(np.asarray(ef.rows), enc) and ef.ids == ids
    if n:
        assert isinstance(ef.rows, np.memmap)
    parts = list(ef.chunks(1000))
    assert sum(c.shape[0] for c in parts) == n
    # our own files read back the same way
    q = tmp_path / "ours.pkl"
    write_embedding_file(str(q), enc, ids)
    rows, names = read_embedding_file(str(q))
    assert np.array_equal(np.asarray(rows), enc) and names == ids


def test_unknown_layout_falls_back_to_pickle(tmp_path):
    enc = np.arange(12, dtype=np.float64).reshape(3, 4)  # float64: not the layout the fast path accepts
    p = tmp_path / "odd.pkl"
    _reference_dump(p, enc, ["a", "b", "c"])
    ef = EmbeddingFile(str(p))
    assert not ef.streaming and ef.rows.dtype == np.float32 and np.array_equal(ef.rows, enc.astype(np.float32))
    assert ef.ids == ["a", "b", "c"]


def test_split_matches_the_reference_script(tmp_path):
    rng = np.random.default_rng(0)
    enc = rng.standard_normal((1001, 24), dtype=np.float32)
    ids = ["d%d" % i for i in range(1001)]
    src = tmp_path / "embeddings.corpus.rank.0"
    _reference_dump(src, enc, ids)
    outs = split_embedding_file(str(src), str(tmp_path / "split"), num_splits=3)
    assert [o.rsplit(".", 1)[1] for o in outs] == ["0", "1", "2"]
    for s, o in enumerate(outs):
        with open(o, "rb") as f:
            e, names = pickle.load(f)
        # scripts/split_embeddings.py:17-21: embedding[split::num_splits], lookup[split::num_splits].tolist()
        assert np.array_equal(e, enc[s::3]) and names == np.array(ids)[s::3].tolist()
