"""Stress the generation-3 SAN path on ONE GPU under SM contention and clock perturbation: full train steps (eager, no graph)
while a second stream keeps a variable number of SMs busy with unrelated kernels (what NCCL kernels do to the step at N > 1).
    gpurun -- 'python scripts/stress_gen3.py [steps] [mode]'      mode: none | matmul | spin"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    mode = sys.argv[2] if len(sys.argv) > 2 else "matmul"
    import bench
    dev = torch.device("cuda", 0)
    model, args, _ = bench.build_model(dev, "bf16")
    model.train()
    from iisan_b200.optim import FusedAdam, param_groups
    from iisan_b200.engine import TrainStep
    from iisan_b200 import _lib
    lib = _lib.load()
    opt = FusedAdam(param_groups(model, args))
    eager = TrainStep(model, opt, use_graph=False)
    if "timing" in mode:
        lib.iisan_timing_enable(1)
    g = torch.Generator(device=dev).manual_seed(7)
    batches = bench.make_device_batches(3, 512, dev, torch.bfloat16, g)
    side = torch.cuda.Stream()
    a = torch.randn(2048, 2048, device=dev, dtype=torch.bfloat16)
    small = torch.randn(256, 256, device=dev)
    t0 = time.time()
    for i in range(steps):
        if "matmul" in mode or "spin" in mode:
            with torch.cuda.stream(side):
                if "matmul" in mode:
                    for _ in range(2 + i % 5):
                        a @ a
                else:
                    for _ in range(20 + i % 30):
                        small.add_(1.0)
        ids, image, text, lm = batches[i % 3]
        if "adam" in mode:
            loss = eager(ids, image, text, lm)
        else:
            model.zero_grad(set_to_none=True)
            loss = model(ids, image, text, lm, dev)
            loss.backward()
        if i % 50 == 0:
            torch.cuda.synchronize()
            print(i, float(loss), flush=True)
    torch.cuda.synchronize()
    print("ok", steps, mode, round(time.time() - t0, 1), "s")


if __name__ == "__main__":
    main()
