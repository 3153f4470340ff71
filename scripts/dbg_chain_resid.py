"""Debug probe for the EXPERIMENTAL TMEM-resident residual of the chain forward (IISAN_B200_CHAIN_TMEM_RESID): compares the
embeddings of the flagged kernel with the default kernel for several item counts / resident-chunk counts, with the workspace
poisoned (all bytes 0xFF = bf16 NaN) so that a read of anything this launch did not write shows up as NaN.

    python scripts/dbg_chain_resid.py            # JSON lines
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch import nn


def main():
    from iisan_b200 import _lib, model as pkg, ops
    from iisan_b200.config import default_args
    from iisan_b200.precision import set_compute_mode
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    args = default_args()

    class ImgStub(nn.Module):
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(768, args.embedding_dim)

    m = pkg.ModelMM(args, 1000, True, ImgStub(), nn.Identity(), [1.0] * 1001)
    m.mm_encoder = pkg.IISANAdaptedMModel(m.mm_encoder, args)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias") or "side_gate" in n:
                p.add_(0.05 * torch.randn_like(p))
    m = m.to(dev)
    set_compute_mode("bf16")
    san = m.mm_encoder
    E = args.embedding_dim
    ops._workspace = lambda nbytes, device: torch.full((max(int(nbytes), 256),), 0xFF, dtype=torch.uint8, device=device)   # poison
    FWD = "IISAN_B200_CHAIN_TMEM_RESID"
    gen = torch.Generator(device=dev).manual_seed(2)
    for N in (44, 176, 5632):
        img = torch.randn(N, 13, 768, device=dev, generator=gen).bfloat16()
        txt = torch.randn(N, 13, 768, device=dev, generator=gen).bfloat16()
        os.environ.pop(FWD, None)
        with torch.no_grad():
            ref = san.embed(img, txt).clone()
            ref2 = san.embed(img, txt).clone()
        rec = {"N": N, "default_nan": bool(torch.isnan(ref).any()), "default_rerun_equal": bool(torch.equal(ref, ref2)), "runs": []}
        for chunks in (10, 8, 4, 1):
            os.environ[FWD] = str(chunks)
            for rep in range(2):
                with torch.no_grad():
                    out = san.embed(img, txt)
                torch.cuda.synchronize()
                d = (out - ref).abs()
                bad_rows = (d.max(dim=1).values > 0).nonzero().flatten()
                rec["runs"].append({"chunks": chunks, "rep": rep, "nan": bool(torch.isnan(out).any()), "equal": bool(torch.equal(out, ref)),
                                    "max_abs": [float(d[:, :E].max()), float(d[:, E:2 * E].max()), float(d[:, 2 * E:].max())],
                                    "ref_max": float(ref.abs().max()), "n_bad_rows": int(bad_rows.numel()),
                                    "bad_row_range": [int(bad_rows.min()), int(bad_rows.max())] if bad_rows.numel() else None})
        os.environ.pop(FWD, None)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
