"""Device timings (CUDA events, median of reps, L2 flushed between reps) of two pieces the whole-step bench does not isolate:

  * iisan_gather_states at the store shape of BASELINE configs[1] (catalogue 19,247 x 7 x 768 bf16 per modality, 5632 ids);
  * the fast in-batch CE (forward + backward) of one rank against a pool of W x 512 users, W = 1, 2, 4, 8 -- the global negative
    pool of a data-parallel step, timed on ONE GPU (the kernels do not care where the pool came from).
    Both split policies of ce_fast_splits are timed (IISAN_B200_CE_ONE_WAVE, read by the library per call).

    python scripts/ce_gather_bench.py            -> one JSON line
"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps, flush):
    out = []
    for _ in range(reps + 2):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(out[2:]), 1)


def main():
    from iisan_b200 import _lib, ops
    dev = torch.device("cuda:0")
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)          # 256 MB > L2
    res = {}
    g = torch.Generator(device=dev).manual_seed(1)
    # ---- gather ----
    items, A, d, n = 19247, 7, 768, 5632
    table = torch.randn(items, A, d, device=dev, generator=g).bfloat16()
    ids = torch.randint(1, items, (n,), device=dev, generator=g)
    sel = torch.arange(A, dtype=torch.int32, device=dev)
    us = timed(lambda: ops.gather_states(table, ids, sel), 10, flush)
    byt = 2 * n * A * d * 2
    res["gather_states"] = {"us": us, "bytes_read_plus_written": byt, "gbs": round(byt / (us * 1e-6) / 1e9)}
    ref = table[ids]
    assert torch.equal(ops.gather_states(table, ids, sel), ref)
    # ---- CE against W x 512 users ----
    B, L, E = 512, 10, 64
    res["ce"] = {}
    for policy, W in [(p, w) for w in (1, 2, 4, 8) for p in ("one_wave", "multi_wave")]:
        os.environ["IISAN_B200_CE_ONE_WAVE"] = "1" if policy == "one_wave" else "0"
        Bc = B * W
        ids_all = torch.randint(1, 22785, (Bc, L + 1), device=dev, generator=g)
        lm_all = torch.ones(Bc, L, device=dev)
        pop = torch.rand(22786, device=dev, generator=g) * 0.9 + 0.05
        prec = (torch.randn(B * L, E, device=dev, generator=g) * 0.7).requires_grad_(True)
        score = (torch.randn(Bc * (L + 1), E, device=dev, generator=g) * 0.5).requires_grad_(True)
        off = B * (W - 1)

        def fwd_bwd():
            prec.grad = None; score.grad = None
            _, _, loss = ops.InBatchCeFn.apply(prec, score, ids_all[off:off + B], ids_all, lm_all[off:off + B], lm_all, pop, off, _lib.COMPUTE_BF16)
            loss.backward()
        # device time of the CE kernel class: the library's own events around every launch (bench.py's kernel_classes)
        import ctypes as C
        lib = _lib.load()
        for _ in range(2):
            fwd_bwd()
        torch.cuda.synchronize()
        k_ce = list(_lib.KERNEL_CLASSES).index("ce")
        tot, cnt = C.c_double(0), C.c_int64(0)
        lib.iisan_timing_enable(1)
        reps = 6
        for _ in range(reps):
            flush.add_(1)
            fwd_bwd()
        torch.cuda.synchronize()
        lib.iisan_timing_enable(0)
        lib.iisan_timing_read(k_ce, C.byref(tot), C.byref(cnt))
        res["ce"][f"W{W}_{policy}"] = {"kernels_us": round(tot.value * 1e3 / reps, 1), "launches": cnt.value // reps,
                              "fwd_bwd_wall_us": timed(fwd_bwd, 6, flush)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
