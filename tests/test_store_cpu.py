"""Host side of the cached-state store: the one-time repack of the reference's per-item .pt files (same file naming and
tensor layout as Code_Cached/preprocess_vectors.py:27-31 / data_utils/dataset.py:29-34) and the id / log_mask batch layout
of Build_MM_Dataset.__getitem__ (dataset.py:65-72)."""
import os

import pytest
import torch


def test_load_state_files_round_trip(tmp_path):
    from iisan_b200.store import load_state_files
    item_num = 9
    keys = {i: f"B00{i:05d}".encode() for i in range(1, item_num + 1)}
    d = tmp_path / "bert_outputs"
    os.makedirs(d)
    ref = {}
    g = torch.Generator().manual_seed(0)
    for i, k in keys.items():
        t = torch.randn(13, 32, generator=g)
        ref[i] = t
        torch.save(t, d / f"bert_{k.decode()}.pt")               # what save_outputs writes
    table = load_state_files(str(d), keys, item_num, "bert", dtype=torch.float32, workers=4)
    assert table.shape == (item_num + 1, 13, 32) and not table[0].any()
    for i in range(1, item_num + 1):
        assert torch.equal(table[i], ref[i])
    tb = load_state_files(str(d), keys, item_num, "bert")         # default bf16 table
    assert tb.dtype == torch.bfloat16 and torch.equal(tb[3], ref[3].bfloat16())
    os.remove(d / f"bert_{keys[4].decode()}.pt")
    with pytest.raises(FileNotFoundError):
        load_state_files(str(d), keys, item_num, "bert")


def test_build_id_batch_matches_reference_layout():
    from iisan_b200.store import build_id_batch
    ids, lm = build_id_batch([[3, 4, 5], list(range(1, 12)), [7, 8]], 10)
    # dataset.py:69-72: mask_len = 11 - len(seq); log_mask = [0]*mask_len + [1]*(len(seq)-1); ids = [0]*mask_len + seq
    assert ids[0].tolist() == [0] * 8 + [3, 4, 5] and lm[0].tolist() == [0.0] * 8 + [1.0, 1.0]
    assert ids[1].tolist() == list(range(1, 12)) and lm[1].tolist() == [1.0] * 10
    assert ids[2].tolist() == [0] * 9 + [7, 8] and lm[2].tolist() == [0.0] * 9 + [1.0]


def test_store_batch_equals_reference_dataset_fixture(tmp_path):
    """tests/golden/data_batch.npz is a train batch produced by the REFERENCE's own ``Build_MM_Dataset.__getitem__`` +
    ``load_output`` (dataset.py:29-34, 65-92) from cache files written by its own ``save_outputs`` (preprocess_vectors.py:27-31),
    default-collated and reshaped as in run.py:368-377 (oracle/make_golden_dataset.py executes that source text).  The product's
    host side -- ``load_state_files`` (one-time repack of the same files), ``build_id_batch`` (ids / log_mask) and the
    gather-by-id semantics of the store (row 0 = zero padding item; the device kernel iisan_gather_states is checked against
    the same indexing in tests/test_gpu_parity.py::test_gather_states_bit_exact) -- must reproduce it bit for bit."""
    import numpy as np
    from iisan_b200.store import build_id_batch, load_state_files
    from oracle.synthetic import make_cache_case
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_batch.npz"))
    keys, bert, vit, u2seq = make_cache_case(int(z["seed"]), layers=int(z["layers"]), d=int(z["d"]))
    for sub, prefix, states in (("bert_outputs", "bert", bert), ("vit_outputs", "vit", vit)):
        os.makedirs(tmp_path / sub)
        for i, t in states.items():
            torch.save(t, tmp_path / sub / f"{prefix}_{keys[i].decode()}.pt")
    item_num = len(keys)
    text_table = load_state_files(str(tmp_path / "bert_outputs"), keys, item_num, "bert", dtype=torch.float32, workers=2)
    image_table = load_state_files(str(tmp_path / "vit_outputs"), keys, item_num, "vit", dtype=torch.float32, workers=2)
    ids, log_mask = build_id_batch([u2seq[u] for u in range(len(u2seq))], 10)
    assert ids.dtype == torch.int64 and log_mask.dtype == torch.float32
    assert np.array_equal(ids.numpy(), z["ids"]) and np.array_equal(ids.view(-1).numpy(), z["flat_ids"])
    assert np.array_equal(log_mask.numpy(), z["log_mask"])
    assert np.array_equal(image_table[ids].numpy(), z["image"])          # [B, 11, 13, d]: padded slots are the zero item
    assert np.array_equal(text_table[ids].numpy(), z["text"])
    sel = [0, 2, 4, 6, 8, 10, 12]                                        # the packed store keeps the selected layers only
    assert np.array_equal(image_table[:, sel][ids].numpy(), z["image"][:, :, sel])
