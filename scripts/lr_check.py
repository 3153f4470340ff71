"""Third generation (resident-state chain forward + low-rank adjoint backward, san_lr.cu) against the second generation on the
same inputs: embeddings, every parameter gradient (relative L2), kernel-class times.
    gpurun -- 'python scripts/lr_check.py [B] [scale]'"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    N = int(sys.argv[3]) if len(sys.argv) > 3 else B * 11
    import bench
    from iisan_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    model, args, _ = bench.build_model(dev, "bf16")
    model.eval()
    san = model.mm_encoder
    g = torch.Generator(device=dev).manual_seed(5)
    with torch.no_grad():                       # gates away from 0.5, biases non-zero, adapters away from their 0.01 init
        for n, p in san.named_parameters():
            if "side_gate" in n:
                p.copy_((torch.rand(p.shape, device=dev, generator=g) - 0.5) * 0.3)
            elif "adapter_list" in n and n.endswith("weight"):
                p.mul_(scale)
            elif n.endswith("bias"):
                p.add_(torch.randn(p.shape, device=dev, generator=g) * 0.05)
    img = torch.randn(N, 13, 768, device=dev, generator=g).bfloat16()
    txt = torch.randn(N, 13, 768, device=dev, generator=g).bfloat16()
    w = torch.randn(N, 192, device=dev, generator=g)
    res = {}
    gens = tuple(int(x) for x in os.environ.get('LR_GENS', '2,3').split(','))
    reps = int(os.environ.get('LR_REPS', '10'))
    for gen in gens:
        lib.iisan_debug_chain_generation(gen)
        san.zero_grad(set_to_none=True)
        out = san.embed(img, txt)
        (out * w).sum().backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().clone() for n, p in san.named_parameters() if p.grad is not None}
        lib.iisan_timing_enable(1)
        for _ in range(reps):
            san.zero_grad(set_to_none=True)
            o = san.embed(img, txt)
            (o * w).sum().backward()
        torch.cuda.synchronize()
        lib.iisan_timing_enable(0)
        t = {}
        for k, name in enumerate(_lib.KERNEL_CLASSES):
            tot, n = C.c_double(0), C.c_int64(0)
            lib.iisan_timing_read(k, C.byref(tot), C.byref(n))
            if n.value:
                t[name] = {"us_per_launch": round(tot.value / n.value * 1e3, 2), "us_per_step": round(tot.value / reps * 1e3, 1), "launches_per_step": n.value / reps}
        res[gen] = (out.detach().clone(), grads, t)
    if len(gens) < 2:
        print(json.dumps(res[gens[0]][2], indent=1)); return
    o2, g2, t2 = res[2]; o3, g3, t3 = res[3]
    rel = {n: float((g2[n] - g3[n]).norm() / (g2[n].norm() + 1e-30)) for n in g2}
    worst = sorted(rel.items(), key=lambda kv: -kv[1])[:12]
    gates = {n: (float(g2[n].reshape(-1)[0]), float(g3[n].reshape(-1)[0])) for n in g2 if "side_gate" in n}
    print(json.dumps({"B": B, "N": N, "scale": scale, "finite": bool(torch.isfinite(o3).all()) and all(bool(torch.isfinite(v).all()) for v in g3.values()),
                      "emb_rel_l2": float((o2 - o3).norm() / o2.norm()), "emb_max_abs_diff": float((o2 - o3).abs().max()),
                      "emb_max_abs": float(o2.abs().max()), "missing_grads": [n for n in g2 if n not in g3],
                      "worst_grad_rel_l2": worst, "gates_gen2_gen3": gates, "timing_gen2": t2, "timing_gen3": t3}, indent=1))


if __name__ == "__main__":
    main()
