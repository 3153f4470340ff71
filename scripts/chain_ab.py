"""A/B of the fused chain kernels: generation 1 vs generation 2 on the same inputs (bit-equality of the embeddings, gradients,
kernel time from CUDA events around the SAN forward / backward).   gpurun -- 'python scripts/chain_ab.py [B] [d]'"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    import bench
    from iisan_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    model, args, _ = bench.build_model(dev, "bf16")
    model.eval()
    san = model.mm_encoder
    N = B * 11
    g = torch.Generator(device=dev).manual_seed(5)
    img = torch.randn(N, 13, 768, device=dev, generator=g).bfloat16()
    txt = torch.randn(N, 13, 768, device=dev, generator=g).bfloat16()
    w = torch.randn(N, 192, device=dev, generator=g)
    res = {}
    for gen in (1, 2):
        lib.iisan_debug_chain_generation(gen)
        san.zero_grad(set_to_none=True)
        out = san.embed(img, txt)
        (out * w).sum().backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().clone() for n, p in san.named_parameters() if p.grad is not None}
        # timing: the chain kernels through the library's per-class event timers
        lib.iisan_timing_enable(1)
        for _ in range(10):
            san.zero_grad(set_to_none=True)
            o = san.embed(img, txt)
            (o * w).sum().backward()
        torch.cuda.synchronize()
        lib.iisan_timing_enable(0)
        import ctypes as C
        t = {}
        for k, name in enumerate(_lib.KERNEL_CLASSES):
            tot, n = C.c_double(0), C.c_int64(0)
            lib.iisan_timing_read(k, C.byref(tot), C.byref(n))
            if n.value:
                t[name] = tot.value / n.value * 1e3
        res[gen] = (out.detach().clone(), grads, t)
    o1, g1, t1 = res[1]; o2, g2, t2 = res[2]
    worst = max(float((g1[n] - g2[n]).norm() / (g1[n].norm() + 1e-30)) for n in g1)
    print(json.dumps({"B": B, "finite": bool(torch.isfinite(o2).all()), "embeddings_bit_equal": bool(torch.equal(o1, o2)),
                      "max_abs_diff": float((o1 - o2).abs().max()), "worst_grad_rel_l2": worst,
                      "us_per_launch_gen1": t1, "us_per_launch_gen2": t2}))


if __name__ == "__main__":
    main()
