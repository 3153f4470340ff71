"""The drop-in boundary under the wrappers the reference's run.py puts around the model (SURVEY.md 8b "Wrappers it must
survive"): `.to(local_rank)`, SyncBatchNorm conversion (run.py:140), the requires_grad routing by name (:185-187),
DDP(find_unused_parameters=False) (:258), torch.optim.Adam over the name-routed learning-rate groups (:260-307), and the
batch loop `zero_grad / autocast() / scaler.scale(loss).backward() / scaler.step / scaler.update` (:368-385).

The body below is that loop with the reference's statements in the reference's order; the model classes come from
iisan_b200.model.  Checked against the oracle trained on the CPU with torch.optim.Adam over the same groups: the loss of
every step within north_star's fast-mode tolerance (autocast selects IISAN_COMPUTE_BF16), every parameter receives a
finite gradient (DDP's contract), the GradScaler neither skips a step nor changes the result.  (CPU stand-in run with the
rounding-point emulation of tests/bf16_emulation.py instead of the kernels: losses within 4.2e-4, update cosines >= 0.99.)
"""
import argparse
import os

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

LRS = dict(lr=2e-4, adapter_cv_lr=1e-4, adapter_bert_lr=1e-4, fine_tune_lr_image=1e-4, fine_tune_lr_text=5e-5)   # CC/scripts/run_IISAN.py:30-43


def _oracle_losses(cfg, params, pop, batches, steps):
    """fp32 CPU: oracle forward/backward + torch.optim.Adam over the reference's groups (oracle-side routing restated by name)."""
    from iisan_b200.optim import param_groups
    from oracle import iisan_oracle as O
    P = O.params_to_torch(params)

    class Named:                                   # param_groups only needs named_parameters()
        def named_parameters(self):
            return list(P.items())

    opt = torch.optim.Adam(param_groups(Named(), argparse.Namespace(**LRS)))
    losses = []
    for s in range(steps):
        opt.zero_grad()
        out = O.model_forward(P, batches[s % len(batches)], pop, cfg)
        out["loss"].backward()
        opt.step()
        losses.append(float(out["loss"].detach()))
    return losses, {k: v.detach().numpy() for k, v in P.items()}


def _reference_style_loop(model, optimizer, batches, steps, local_rank, use_scaler, autocast_dtype):
    scaler = torch.cuda.amp.GradScaler() if use_scaler else None
    losses = []
    model.train()
    for s in range(steps):
        b = batches[s % len(batches)]
        sample_items_id = torch.from_numpy(b["ids"]); sample_items_image = torch.from_numpy(b["image"])
        sample_items_text = torch.from_numpy(b["text"]); log_mask = torch.from_numpy(b["log_mask"])
        sample_items_id, sample_items_image, sample_items_text, log_mask = \
            sample_items_id.to(local_rank), sample_items_image.to(local_rank), sample_items_text.to(local_rank), log_mask.to(local_rank)
        sample_items_image = sample_items_image.view(-1, 11, 13, 768)
        sample_items_text = sample_items_text.view(-1, 11, 13, 768)
        sample_items_id = sample_items_id.view(-1)
        optimizer.zero_grad()
        with torch.cuda.amp.autocast(dtype=autocast_dtype):
            bz_loss = model(sample_items_id, sample_items_image, sample_items_text, log_mask, local_rank)
        if scaler is not None:
            scaler.scale(bz_loss).backward()
            scaler.step(optimizer)
            scaler.update()
        else:
            bz_loss.backward()
            optimizer.step()
        losses.append(float(bz_loss.data.float()))
    return losses, scaler


@pytest.fixture(scope="module")
def process_group():
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        import socket
        with socket.socket() as sock:                    # a free port: the box may run other rendezvous at the same time
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
        created = True
    yield
    if created:
        dist.destroy_process_group()


STEPS = 4


@pytest.fixture(scope="module")
def workload():
    """Seeded parameters, one distinct batch per step (a repeated batch is memorised within two Adam steps, which turns the
    comparison into a test of chaotic amplification), and the oracle's fp32 training run."""
    from oracle.synthetic import PathConfig, make_batch, make_params, make_pop_prob
    cfg = PathConfig(item_num=300)
    params = make_params(cfg, 31, perturb=True)
    pop = make_pop_prob(cfg, 31)
    batches = [make_batch(16, cfg, 40 + i, "realistic") for i in range(STEPS)]
    ref_losses, ref_params = _oracle_losses(cfg, params, pop, batches, STEPS)
    return cfg, params, pop, batches, ref_losses, ref_params


@pytest.mark.parametrize("use_scaler,autocast_dtype", [(True, torch.float16), (False, torch.bfloat16)])
def test_reference_batch_loop_with_ddp_autocast_gradscaler(process_group, workload, use_scaler, autocast_dtype):
    from torch.nn.parallel import DistributedDataParallel as DDP
    from iisan_b200 import model as pkg
    from iisan_b200.optim import param_groups
    from oracle.synthetic import make_args
    local_rank = 0
    torch.cuda.set_device(local_rank)
    cfg, params, pop, batches, ref_losses, ref_params = workload
    args = make_args(cfg)
    for k, v in LRS.items():
        setattr(args, k, v)
    steps = STEPS

    class ImgStub(nn.Module):                       # ViTForImageClassification stand-in
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(768, cfg.embedding_dim)

    # ---- Code_Cached/run.py:138-140, 182-187, 258 ----
    model = pkg.ModelMM(args, cfg.item_num, True, ImgStub(), nn.Identity(), pop).to(local_rank)
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    for name, param in model.named_parameters():
        param.requires_grad = False                 # fine_tune_to == "None": run.py:156-158
    model.mm_encoder = pkg.IISANAdaptedMModel(model.mm_encoder, args).to(local_rank)
    for index, (name, param) in enumerate(model.named_parameters()):
        if any(["user" in name, "classifier" in name, "title.fc" in name, "cv_pre_fc" in name, "bert_pre_fc" in name]) or \
                all(["user" not in name, "encoder" not in name]):
            param.requires_grad = True
    assert [n for n, _ in model.named_parameters()] == list(params.keys())
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(torch.from_numpy(params[n]))
    assert all(p.requires_grad for p in model.parameters())        # every tensor of the cached path trains
    model = DDP(model, device_ids=[local_rank], output_device=local_rank, find_unused_parameters=False)
    optimizer = torch.optim.Adam(param_groups(model.module, args))   # run.py:260-307
    assert sum(len(g["params"]) for g in optimizer.param_groups) == 146

    # dropout off for the comparison (the SAN has none; SASRec's drop_rate is 0 in PathConfig) but the loop calls model.train()
    losses, scaler = _reference_style_loop(model, optimizer, batches, steps, local_rank, use_scaler, autocast_dtype)
    print("losses", losses, "oracle", ref_losses, "scale", None if scaler is None else scaler.get_scale())
    assert all(np.isfinite(losses))
    for got, ref in zip(losses, ref_losses):
        assert abs(got - ref) <= 1e-2 * abs(ref), (losses, ref_losses)
    if scaler is not None:
        assert scaler.get_scale() == 65536.0        # no inf/nan was found: no step skipped, no back-off
    for n, p in model.module.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # the parameters moved the way the oracle's did: Adam normalises every coordinate to ~lr per step, so compare the
    # accumulated update of the large tensors by cosine (coordinates with a gradient inside the bf16 noise may flip sign)
    for n, p in model.module.named_parameters():
        if p.numel() < 4096:
            continue
        du = (p.detach().cpu().numpy() - params[n]).ravel().astype(np.float64)
        dr = (ref_params[n] - params[n]).ravel().astype(np.float64)
        cos = float(du @ dr / (np.linalg.norm(du) * np.linalg.norm(dr) + 1e-30))
        assert cos >= 0.8, (n, cos)
