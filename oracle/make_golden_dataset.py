"""TEST INFRASTRUCTURE -- freeze one train batch produced by the REFERENCE's own data loader -> tests/golden/data_batch.npz.

``Code_Cached/data_utils/dataset.py`` cannot be imported here (``import lmdb`` at its top, and lmdb is not installed), so the
source text of ``load_output`` / ``Build_MM_Dataset`` (dataset.py:29-34, 36-92) and of the writer ``save_outputs``
(Code_Cached/preprocess_vectors.py:27-31) is read from /root/reference and executed as is (nothing is copied into the
repository).  ``Build_MM_Dataset.__init__`` opens the image lmdb, which the cached path never reads afterwards: the instance is
created without it and given exactly the attributes ``__getitem__`` uses.  The batch is then collated and reshaped the way
the training loop does it (Code_Cached/run.py:134-135 DataLoader default collate, :368-377 ``view(-1, 11, 13, D)`` /
``view(-1)``).  Run in the build container only:

    python -m oracle.make_golden_dataset
"""
import ast
import os
import tempfile

import numpy as np
import torch
from torch.utils.data import Dataset
from torch.utils.data.dataloader import default_collate

REF_DS = "/root/reference/Code_Cached/data_utils/dataset.py"
REF_PRE = "/root/reference/Code_Cached/preprocess_vectors.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "data_batch.npz")


def _exec_defs(path, names):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(body) == len(names), (path, names)
    ns = {"torch": torch, "np": np, "os": os, "Dataset": Dataset}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    import sys
    sys.path.insert(0, ROOT)
    from oracle.synthetic import make_cache_case
    seed, layers, d = 4242, 13, 16
    keys, bert, vit, u2seq = make_cache_case(seed, layers=layers, d=d)
    save_outputs = _exec_defs(REF_PRE, ["save_outputs"])["save_outputs"]
    ns = _exec_defs(REF_DS, ["load_output", "Build_MM_Dataset"])
    with tempfile.TemporaryDirectory() as tmp:
        # the reference's writer, keyed by the decoded ASIN like preprocess_vectors.py:89-103 does
        save_outputs(os.path.join(tmp, "bert_outputs"), {keys[i].decode(): t for i, t in bert.items()}, prefix="bert")
        save_outputs(os.path.join(tmp, "vit_outputs"), {keys[i].decode(): t for i, t in vit.items()}, prefix="vit")
        ds = object.__new__(ns["Build_MM_Dataset"])
        ds.u2seq, ds.item_id_to_keys, ds.max_seq_len, ds.stored_vector_path = u2seq, keys, 10 + 1, tmp     # dataset.py:38-60
        samples = [ds[u] for u in range(len(ds))]
    ids, image, text, log_mask = default_collate(samples)
    image = image.view(-1, 11, layers, d); text = text.view(-1, 11, layers, d); flat_ids = ids.view(-1)     # run.py:373-377
    np.savez_compressed(OUT, ids=ids.numpy(), flat_ids=flat_ids.numpy(), image=image.numpy(), text=text.numpy(),
                        log_mask=log_mask.numpy(), seed=seed, layers=layers, d=d)
    print("wrote", OUT, tuple(ids.shape), tuple(image.shape), ids.dtype, image.dtype, log_mask.dtype)


if __name__ == "__main__":
    main()
