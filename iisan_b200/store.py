"""HBM-resident cached hidden-state store (the cached-state path of BASELINE.json north_star (1)).

The reference re-loads one ``<prefix>_<ASIN>.pt`` file per item and modality for every sample of every step
(Code_Cached/data_utils/dataset.py:29-34, 65-92) and ships all 13 layers of every slot over the host link (run.py:370-374).
Here the catalogue is packed ONCE into one table per modality that keeps only the layers the towers read,

    table[item_id, k, :] = states[item_id][sel[k], :]        (row 0 = the all-zero padding item, dataset.py:87-88)

(Instrument: 19,247 x 7 x 768 bf16 = 207 MB per modality), resident in HBM.  A train batch is then just the item ids: the
per-item, per-layer gather runs on the device (iisan_gather_states: 128-bit loads, padding ids zero-filled without touching
the table) and the model consumes the packed [N, A, d] tensors directly (``packed=True``).
"""
from __future__ import annotations

import torch

from . import ops


class CachedStateStore:
    def __init__(self, image_states, text_states, sel_img, sel_text, device="cuda", dtype=torch.bfloat16):
        """``image_states`` / ``text_states``: [item_num + 1, layers, d] tensors (any device; row 0 is ignored and treated as the
        zero padding item) -- e.g. the stacked contents of stored_vectors_*/{vit,bert}_<ASIN>.pt in item-id order."""
        self.sel_img = sorted(set(int(v) for v in sel_img))
        self.sel_text = sorted(set(int(v) for v in sel_text))
        self.image = self._pack(image_states, self.sel_img, device, dtype)
        self.text = self._pack(text_states, self.sel_text, device, dtype)
        self._all_img = torch.arange(len(self.sel_img), dtype=torch.int32, device=device)
        self._all_text = torch.arange(len(self.sel_text), dtype=torch.int32, device=device)

    @classmethod
    def for_model(cls, model, image_states, text_states, **kw):
        plan = model.mm_encoder.plan
        return cls(image_states, text_states, plan.layers_img_sel, plan.layers_text_sel, **kw)

    @staticmethod
    def _pack(states, sel, device, dtype):
        out = torch.empty(states.shape[0], len(sel), states.shape[2], dtype=dtype, device=device)
        step = max(1, (256 << 20) // max(1, states[0].numel() * states.element_size()))        # ~256 MB of source per slice
        idx = torch.as_tensor(sel)
        for lo in range(0, states.shape[0], step):
            out[lo:lo + step] = states[lo:lo + step].index_select(1, idx.to(states.device)).to(device=device, dtype=dtype)
        out[0].zero_()
        return out

    def gather(self, ids):
        """ids int64 [n] on the store's device -> (image [n, A_i, d], text [n, A_t, d]); id 0 -> zeros (bit-exact selection)."""
        return ops.gather_states(self.image, ids, self._all_img), ops.gather_states(self.text, ids, self._all_text)
