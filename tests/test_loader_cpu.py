"""pb_index_load (MmapIndex::load, index.rs:1026-1139) validates the reference's index directory before it touches the
device, so the error contract is checked here without a GPU: malformed or unsupported directories are refused with the
status the header states, and a well-formed one gets as far as the device (PB_ERR_CUDA on a box without one)."""
import json
import os
import shutil

import numpy as np
import pytest


@pytest.fixture(scope="module")
def npb():
    import next_plaid_b200 as m
    m.build_library()
    return m


@pytest.fixture(scope="module")
def index_dir(oracle, tmp_path_factory):
    docs = oracle.synthetic_corpus(120, 12, dim=32, seed=3, ragged=True)
    ix = oracle.create_index(docs, nbits=4, seed=1, num_partitions=16)
    d = tmp_path_factory.mktemp("ix")
    oracle.write_index(ix, str(d), chunk_docs=50)          # three chunks
    return str(d), ix


def _copy(src, tmp_path):
    dst = tmp_path / "ix"
    shutil.copytree(src, dst)
    return str(dst)


def _status(npb, path):
    with pytest.raises(npb.PlaidError) as e:
        npb.MmapIndex.load(path)
    return e.value.status, str(e.value)


def test_well_formed_directory_reaches_the_device(npb, index_dir):
    path, ix = index_dir
    if npb.device_count() > 0:
        gpu = npb.MmapIndex.load(path)
        assert gpu.num_documents() == ix.num_documents and gpu.num_embeddings() == ix.num_embeddings
        gpu.close()
    else:
        st, msg = _status(npb, path)
        assert st == 2, msg                                  # PB_ERR_CUDA: parsed fine, no device, no fallback


def _as_fast_plaid(path):
    """Rewrite a next-plaid directory with fast-plaid's dtypes (mmap.rs:1757-1811): <f2 floats, <i8 ivf_lengths."""
    for name in ("centroids.npy", "bucket_weights.npy", "bucket_cutoffs.npy", "avg_residual.npy"):
        p = os.path.join(path, name)
        if os.path.exists(p):
            np.save(p, np.load(p).astype(np.float16))
    p = os.path.join(path, "ivf_lengths.npy")
    np.save(p, np.load(p).astype(np.int64))


def test_fast_plaid_directory_parses(npb, index_dir, tmp_path):
    # read as it is (widened / narrowed in memory), never modified; without a device the load still ends in PB_ERR_CUDA
    path = _copy(index_dir[0], tmp_path)
    _as_fast_plaid(path)
    before = {f: os.path.getmtime(os.path.join(path, f)) for f in os.listdir(path)}
    if npb.device_count() > 0:
        gpu = npb.MmapIndex.load(path)
        assert gpu.num_documents() == index_dir[1].num_documents
        gpu.close()
    else:
        st, msg = _status(npb, path)
        assert st == 2, msg
    assert before == {f: os.path.getmtime(os.path.join(path, f)) for f in os.listdir(path)}
    np.save(os.path.join(path, "ivf_lengths.npy"), np.load(os.path.join(path, "ivf_lengths.npy")).astype(np.float64))
    st, msg = _status(npb, path)
    assert st == 3 and "ivf_lengths" in msg


def test_truncated_payload(npb, index_dir, tmp_path):
    path = _copy(index_dir[0], tmp_path)
    f = os.path.join(path, "1.residuals.npy")
    data = open(f, "rb").read()
    open(f, "wb").write(data[:len(data) - 64])
    st, msg = _status(npb, path)
    assert st == 3 and "truncated" in msg


def test_inconsistent_counts(npb, index_dir, tmp_path):
    path = _copy(index_dir[0], tmp_path)
    lens = np.load(os.path.join(path, "ivf_lengths.npy"))
    lens[0] += 1
    np.save(os.path.join(path, "ivf_lengths.npy"), lens)
    st, msg = _status(npb, path)
    assert st == 3 and "ivf" in msg
    path2 = _copy(index_dir[0], tmp_path / "b")
    meta = json.load(open(os.path.join(path2, "metadata.json")))
    meta["num_embeddings"] += 5
    json.dump(meta, open(os.path.join(path2, "metadata.json"), "w"))
    st, msg = _status(npb, path2)
    assert st == 3 and "num_embeddings" in msg


def test_bad_nbits_and_wrong_dtypes(npb, index_dir, tmp_path):
    path = _copy(index_dir[0], tmp_path)
    meta = json.load(open(os.path.join(path, "metadata.json")))
    meta["nbits"] = 3
    json.dump(meta, open(os.path.join(path, "metadata.json"), "w"))
    st, msg = _status(npb, path)
    assert st == 1 and "divisor of 8" in msg                 # codec.rs:161-166
    path2 = _copy(index_dir[0], tmp_path / "b")
    codes = np.load(os.path.join(path2, "0.codes.npy"))
    np.save(os.path.join(path2, "0.codes.npy"), codes.astype(np.int32))
    st, msg = _status(npb, path2)
    assert st == 3 and "0.codes.npy" in msg                   # chunk files are checked before the device is opened


def test_missing_chunk_file(npb, index_dir, tmp_path):
    path = _copy(index_dir[0], tmp_path)
    os.remove(os.path.join(path, "doclens.2.json"))
    st, msg = _status(npb, path)
    assert st == 3 and "doclens.2.json" in msg
