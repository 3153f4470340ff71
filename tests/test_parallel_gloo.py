"""world_size-2 gloo test (CPU) of the data-parallel host logic: all-gather of the item pool with
reduce-scatter backward, global valid-row normalisation, user offsets.  The loss tile itself is supplied by
the oracle here (test infrastructure); on the GPU the same function runs the CUDA kernel."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_ce(prec, score_all, ids, ids_all, lm, lm_all, pop, user_offset, compute):
    from oracle import iisan_oracle as O
    debias = torch.log(pop[ids_all.reshape(-1)])
    n_valid = int((lm.reshape(-1) != 0).sum())
    loss, _ = O.inbatch_ce(prec, score_all, debias, ids.numpy(), lm.numpy(), ids_all.numpy(), lm_all.numpy(),
                           user_offset=user_offset, n_valid_total=1)
    return loss, torch.tensor([n_valid], dtype=torch.int32)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iisan_b200.parallel import global_negative_loss, allreduce_gradients
    from oracle.synthetic import PathConfig, make_ids, make_pop_prob
    cfg = PathConfig(item_num=60)
    B, E, L = 6, 16, 10
    ids, lm = make_ids(B * world, cfg, 5, "realistic")
    g = torch.Generator().manual_seed(1)
    score_full = torch.randn(B * world * 11, E, generator=g)
    prec_full = torch.randn(B * world * L, E, generator=g)
    pop = torch.from_numpy(make_pop_prob(cfg, 5))
    sl = slice(rank * B, (rank + 1) * B)
    score = score_full[rank * B * 11:(rank + 1) * B * 11].clone().requires_grad_(True)
    prec = prec_full[rank * B * L:(rank + 1) * B * L].clone().requires_grad_(True)
    loss = global_negative_loss(prec, score, torch.from_numpy(ids[sl]).reshape(-1), torch.from_numpy(lm[sl]), pop,
                                ce_fn=_oracle_ce, grad_average=False)
    loss.backward()
    # flat-bucket all-reduce helper: sum of per-rank losses == single-process loss
    w = torch.nn.Parameter(torch.zeros(3)); w.grad = torch.full((3,), float(rank + 1))
    allreduce_gradients([w], average=True)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=loss.detach().numpy(), dscore=score.grad.numpy(),
             dprec=prec.grad.numpy(), w=w.grad.numpy())
    dist.destroy_process_group()


def test_global_negative_pool_two_ranks(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import iisan_oracle as O
    from oracle.synthetic import PathConfig, make_ids, make_pop_prob
    cfg = PathConfig(item_num=60)
    B, E, L = 6, 16, 10
    ids, lm = make_ids(B * world, cfg, 5, "realistic")
    g = torch.Generator().manual_seed(1)
    score = torch.randn(B * world * 11, E, generator=g).requires_grad_(True)
    prec = torch.randn(B * world * L, E, generator=g).requires_grad_(True)
    pop = torch.from_numpy(make_pop_prob(cfg, 5))
    debias = torch.log(pop[torch.from_numpy(ids.reshape(-1))])
    loss, _ = O.inbatch_ce(prec, score, debias, ids, lm, ids, lm)          # single process, concatenated batch
    loss.backward()
    r = [np.load(os.path.join(tmp_path, f"r{k}.npz")) for k in range(world)]
    np.testing.assert_allclose(sum(float(x["loss"]) for x in r), loss.item(), rtol=1e-5)
    np.testing.assert_allclose(np.concatenate([x["dscore"] for x in r]), score.grad.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(np.concatenate([x["dprec"] for x in r]), prec.grad.numpy(), rtol=1e-4, atol=1e-7)
    for x in r:
        np.testing.assert_allclose(x["w"], 1.5)
