"""Build the product model (iisan_b200.model / model_asym) the way Code_Cached/run.py:138,182-183 builds the
reference, and load seeded parameters into it."""
import torch
from torch import nn


def build_product(cfg, params_np, pop, device="cuda"):
    from oracle.synthetic import make_args
    if cfg.asym:
        from iisan_b200 import model_asym as pkg
    else:
        from iisan_b200 import model as pkg
    args = make_args(cfg)

    class ImgStub(nn.Module):
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(cfg.d_img, cfg.embedding_dim)

    m = pkg.ModelMM(args, cfg.item_num, True, ImgStub(), nn.Identity(), pop)
    m.mm_encoder = pkg.IISANAdaptedMModel(m.mm_encoder, args)
    names = [n for n, _ in m.named_parameters()]
    assert names == list(params_np.keys())
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(torch.from_numpy(params_np[n]))
    return m.to(device)


def run_step(model, batch, device="cuda", dtype=torch.float32):
    ids = torch.from_numpy(batch["ids"]).to(device).view(-1)
    image = torch.from_numpy(batch["image"]).to(device=device, dtype=dtype)
    text = torch.from_numpy(batch["text"]).to(device=device, dtype=dtype)
    lm = torch.from_numpy(batch["log_mask"]).to(device)
    model.zero_grad(set_to_none=True)
    loss = model(ids, image, text, lm, device)
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().float().cpu().numpy()) for n, p in model.named_parameters()}
    return loss.detach().float().cpu().numpy(), grads


def fused_chain_route(cfg, plan):
    """Mirror of san_chain_eligible (iisan_b200/csrc/san_bf16.cu) WITHOUT the stored dtype: fp32 / fp16 states of such a
    configuration are packed and rounded to bf16 once (ops.SanFn -> iisan_pack_states) and take the same fused kernels."""
    return bool(cfg.d_text == cfg.d_img and cfg.d_text % 64 == 0 and cfg.d_text >= 640 and cfg.r_cv == 64 and cfg.r_bert == 64 and
                cfg.remove_first != "TRUE" and len(plan) <= 8 and all(None not in st for st in plan) and
                getattr(cfg, "adapter_activation", "RELU") != "GELU")


def emulation_batch(batch, fused, stored_dtype):
    """What the rounding-point emulation must see: on the fused route the product rounds the stored states to bf16 first."""
    import torch
    if not fused or stored_dtype == torch.bfloat16:
        return batch
    r = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    return dict(batch, image=r(batch["image"]), text=r(batch["text"]))
