#!/usr/bin/env python
"""IISAN(Cached) training hot path benchmark (BASELINE.json metric: train samples/s, 1 sample = 1 user = 11 slots).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over one synthetic batch: SAN forward -> fusion -> SASRec -> in-batch CE
-> backward -> Adam.  Workload at N=1 is BASELINE configs[1] (Instrument shape: item_num 19,246, B=512 users,
BERT-base + ViT-B/16 cached states [13, 768] stored bf16, random-init adapters).  For N>1 every rank gets its own
B=512 users (weak scaling, BASELINE configs[2]) with the global in-batch negative pool (item-embedding all-gather).

Prints ONE JSON line (see the keys at the bottom).  `--impl reference` times the reference algorithm's CPU path
(the oracle port -- /root/reference is not present on the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "iisan_cached_train_samples_per_s"
UNIT = "samples/s"
ITEM_NUM = 19246          # Instrument catalogue (BASELINE configs[1]); N > 1 uses the Office catalogue (configs[2])
ITEM_NUM_OFFICE = 22785
SEED = 12345              # reference seed (Code_Cached/scripts/run_IISAN.py:44)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the chain kernels: read from the newest committed ncu summary that
# names them (profiles/*_chain_traffic.json, written by scripts/ncu_traffic.py from an `ncu --set full` capture of this commit's
# kernels); the file name travels into the JSON line.  None when no such file exists.
def ncu_traffic():
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_chain_traffic.json")))
    if not files:
        return {}, None
    try:
        return json.load(open(files[-1])), os.path.relpath(files[-1], ROOT)
    except Exception:
        return {}, None
LRS = dict(lr=2e-4, adapter_cv_lr=1e-4, adapter_bert_lr=1e-4, fine_tune_lr_image=1e-4, fine_tune_lr_text=5e-5)

# Workloads (BASELINE.json configs).  "instrument" (configs[1] / [2]) is the default the driver runs; the IISAN-Versa ones
# (Code_Cached_Asym) are selected with --workload.  layers = cached states per item (n_layers + 1); lists as in the reference's
# launchers (Code_Cached_Asym/script/run_IISAN_eva.py:56-65 for LLaMA-3-70B + EVA-CLIP).
WORKLOADS = {
    "instrument": dict(asym=False, d_text=768, d_img=768, layers_text=13, layers_img=13, vit="1,3,5,7,9,11", bert="1,3,5,7,9,11",
                       stored="bfloat16", batch=512, cpu_batch=512, label="BERT-base+ViT-B/16 cached states [13,768]"),
    "versa_large": dict(asym=True, d_text=1024, d_img=1024, layers_text=25, layers_img=25, vit="1,3,5,7,9,11",
                        bert="1,3,5,7,9,11,13,15,17,19,21,23", stored="bfloat16", batch=512, cpu_batch=64,
                        label="IISAN-Versa BERT-large [25,1024] + ViT-large [25,1024], group layer-drop (13 text / 7 image adapters)"),
    "llama_eva": dict(asym=True, d_text=8192, d_img=5120, layers_text=81, layers_img=49, vit="2,11,20,29,38,47",
                      bert="4,19,34,49,64,79", stored="float16", batch=512, cpu_batch=8,
                      label="IISAN-Versa LLaMA-3-70B-shaped text states [81,8192] + EVA-CLIP image states [49,5120], down_project 8192->5120"),
}


def workload_args(w):
    from iisan_b200.config import default_args
    over = dict(LRS, side_adapter_vit_list=w["vit"], side_adapter_bert_list=w["bert"])
    if w["asym"]:
        over.update(text_embedding_dim=w["d_text"], image_embedding_dim=w["d_img"], text_layers=w["layers_text"] - 1,
                    image_layers=w["layers_img"] - 1, word_embedding_dim=w["d_text"])
    return default_args(**over)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="instrument", choices=["instrument", "versa_large", "llama_eva"],
                    help="instrument = BASELINE configs[1]/[2] (the default the driver runs); versa_large = configs[3]; llama_eva = configs[4]")
    ap.add_argument("--batch", type=int, default=0, help="users per GPU (default: the workload's, 512)")
    ap.add_argument("--compute", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=0, help="users per CPU-arm step (default: the workload's bounded sample)")
    ap.add_argument("--cpu-steps", type=int, default=3, help="cpu_baseline: at least this many timed steps of the CPU arm ...")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="... and at least this much timed CPU work (at most 64 steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-store", action="store_true", help="skip the HBM-resident store e2e measurement")
    ap.add_argument("--no-host-e2e", action="store_true", help="skip the end-to-end measurement from pinned host batches of the reference shapes")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--negatives", default="global", choices=["global", "local"],
                    help="N > 1: in-batch negative pool of the headline value (global = item-embedding all-gather, BASELINE configs[2]; "
                         "local = the reference's DDP semantics); the other mode is measured too and reported under its own key")
    ap.add_argument("--reps", type=int, default=10, help="repetitions of the timed --steps block; the median is reported")
    ap.add_argument("--item-num", type=int, default=0, help="catalogue size (default: Instrument 19,246 at N=1, Office 22,785 at N>1)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_loss(params_np, batch_np, pop, item_num, wname="instrument"):
    """Forward of the oracle on GIVEN parameters / batch (the GPU arm's initial replica and its first batch): the self-check of
    the bench line (loss_ref_step0 vs loss_gpu_step0)."""
    import torch
    from oracle import iisan_oracle as O
    cfg = oracle_config(wname, item_num)
    with torch.no_grad():
        out = O.model_forward(O.params_to_torch(params_np, requires_grad=False), batch_np, pop, cfg)
    return float(out["loss"])


def oracle_config(wname, item_num):
    from oracle.synthetic import PathConfig
    w = WORKLOADS[wname]
    if not w["asym"]:
        return PathConfig(item_num=item_num)
    return PathConfig(item_num=item_num, asym=True, d_text=w["d_text"], d_img=w["d_img"], layers_text=w["layers_text"],
                      layers_img=w["layers_img"], vit_list=w["vit"], bert_list=w["bert"])


def cpu_reference_steps(batch, steps, warmup, wname="instrument", min_seconds=0.0, max_steps=64):
    """fwd + bwd + Adam of the oracle restatement (fp32, all host threads).  Returns (samples/s, s/step, cores, last loss, steps timed).
    ``min_seconds``: keep stepping past ``steps`` until that much CPU work has been timed (at most ``max_steps`` steps)."""
    import numpy as np
    import torch
    from oracle import iisan_oracle as O
    from oracle.synthetic import make_ids, make_params, make_pop_prob
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = oracle_config(wname, ITEM_NUM)
    w = WORKLOADS[wname]
    ids, lm = make_ids(batch, cfg, SEED, "dense")
    g = torch.Generator().manual_seed(SEED)
    image = torch.randn(batch, 11, w["layers_img"], w["d_img"], generator=g).numpy()
    text = torch.randn(batch, 11, w["layers_text"], w["d_text"], generator=g).numpy()
    b = {"ids": ids, "log_mask": lm, "image": image, "text": text}
    pop = make_pop_prob(cfg, SEED)
    P = O.params_to_torch(make_params(cfg, SEED, perturb=False))
    opt = torch.optim.Adam(list(P.values()), lr=LRS["lr"])
    times = []
    it = 0
    while len(times) < steps or (sum(times) < min_seconds and len(times) < max_steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = O.model_forward(P, b, pop, cfg)
        out["loss"].backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
    sps = batch * len(times) / sum(times)
    return sps, sum(times) / len(times), cores, float(out["loss"].item()), len(times)


def workload_name(batch, stored, item_num=ITEM_NUM, wname="instrument"):
    shape = {ITEM_NUM: "Instrument", ITEM_NUM_OFFICE: "Office"}.get(item_num, "custom")
    w = WORKLOADS[wname]
    if wname == "instrument":
        return (f"IISAN(Cached) {shape} shape: item_num {item_num}, B={batch} users/GPU x 11 slots, BERT-base+ViT-B/16 "
                f"cached states [13,768] stored {stored}, 7 of 13 layers, r=64, E=64, random-init adapters, "
                f"dense batch, fwd+bwd+Adam")
    return (f"{w['label']}; {shape} catalogue (item_num {item_num}), B={batch} users/GPU x 11 slots, states stored {stored}, r=64, E=64, "
            f"random-init adapters, dense batch, fwd+bwd+Adam")


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu_batch = a.cpu_batch or WORKLOADS[a.workload]["cpu_batch"]
    a.cpu_batch = cpu_batch
    sps, sec, cores, loss, _n = cpu_reference_steps(cpu_batch, a.steps, a.warmup, a.workload)
    sample = (f"{a.steps} full train steps (fwd+bwd+Adam) of B={a.cpu_batch} dense users, fp32, torch CPU, oracle port of the reference "
              f"algorithm (its negative masks are vectorised; the reference's own per-user Python mask loop, Code_Cached/model/model.py:92-100, "
              f"is slower: SURVEY 8a row a6)")
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a.cpu_batch, "float32", ITEM_NUM, a.workload), "negatives": "local", "step_runner": "torch CPU eager, all host threads",
                   "parallelism": "host cores of rank 0"},
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": loss,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and hence its first-touch / pinned allocations) to the CPUs of the NUMA node the GPU hangs
    off: with one process per GPU the pinned staging buffers otherwise land on an arbitrary socket and the H2D copies of 8 ranks
    share the inter-socket link.  Reads sysfs; silently does nothing when the topology is not exposed."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        if hasattr(pr, "pci_bus_id") and hasattr(pr, "pci_domain_id"):
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:                                                   # e.g. "00000000:1B:00.0"
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = vis.split(",")[local] if vis else str(local)
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", idx], capture_output=True, text=True).stdout
            dom, bus, rest = out.strip().lower().split(":")
            bdf = f"{dom[-4:]}:{bus}:{rest}"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()), len(cpus)
    except Exception:
        pass
    return None


def build_model(device, compute, item_num=ITEM_NUM, wname="instrument"):
    import torch
    from torch import nn
    from iisan_b200.precision import set_compute_mode
    w = WORKLOADS[wname]
    if w["asym"]:
        from iisan_b200 import model_asym as pkg
    else:
        from iisan_b200 import model as pkg
    args = workload_args(w)
    cfg = args
    torch.manual_seed(SEED)

    class ImgStub(nn.Module):                                  # ViTForImageClassification head stand-in (run.py:44-49)
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(w["d_img"], args.embedding_dim)

    import numpy as np
    rng = np.random.default_rng(SEED)
    counts = np.floor(1.0 + rng.pareto(1.2, size=item_num) * 3.0)
    pop = np.concatenate([[1.0], counts / counts.sum()]).astype("float32")
    m = pkg.ModelMM(args, item_num, True, ImgStub(), nn.Identity(), pop)
    m.mm_encoder = pkg.IISANAdaptedMModel(m.mm_encoder, args)          # Code_Cached/run.py:182-183
    set_compute_mode(compute)
    return m.to(device), args, cfg


def make_device_batches(n, B, device, dtype, gen, item_num=ITEM_NUM, wname="instrument"):
    import torch
    w = WORKLOADS[wname]
    out = []

    def states(layers, d):                                     # generated slot by slot: the fp32 staging of [B,11,81,8192] would be 15 GB
        t = torch.empty(B, 11, layers, d, device=device, dtype=dtype)
        for k in range(11):
            t[:, k] = torch.randn(B, layers, d, device=device, generator=gen, dtype=torch.float32).to(dtype)
        return t

    for _ in range(n):
        ids = torch.randint(1, item_num + 1, (B * 11,), device=device, generator=gen, dtype=torch.int64)
        image = states(w["layers_img"], w["d_img"])
        text = states(w["layers_text"], w["d_text"])
        lm = torch.ones(B, 10, device=device, dtype=torch.float32)
        out.append((ids, image, text, lm))
    return out


def run_ours(a):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from iisan_b200 import _lib
    from iisan_b200.engine import PipelinedTrainStep, TrainStep
    from iisan_b200.optim import FusedAdam, param_groups
    lib = _lib.load()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None      # N = 1 keeps every core for the cpu_baseline leg
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    W = WORKLOADS[a.workload]
    item_num = a.item_num or (ITEM_NUM if (world == 1 and a.workload == "instrument") else ITEM_NUM_OFFICE)
    state_dtype = getattr(torch, W["stored"]) if a.compute == "bf16" else torch.float32
    model, args, cfg = build_model(device, a.compute, item_num, a.workload)
    if world > 1:
        for p in model.parameters():                                 # same initial replica on every rank (DDP does this at wrap time)
            dist.broadcast(p.data, 0)
    use_graph = not a.no_graph
    opt = FusedAdam(param_groups(model, args))            # iisan_adam_step: Adam over the reference's 5 LR groups, one launch
    gen = torch.Generator(device=device).manual_seed(SEED + rank)
    B = a.batch or W["batch"]
    n_rot = 3                                               # 3 x 225 MB (bf16) rotating inputs >> 126 MB L2
    batches = make_device_batches(n_rot, B, device, state_dtype, gen, item_num, a.workload)
    group = dist.group.WORLD if world > 1 else None
    dbg = (lambda m: print(f"[bench rank {rank}] {m}", file=sys.stderr, flush=True)) if os.environ.get("IISAN_BENCH_DEBUG") else (lambda m: None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, fn, reps=1):
        """`reps` back-to-back blocks of `n` steps, each block between its own pair of CUDA events; barrier + synchronize on both
        sides of the whole region; per block the MAX over ranks; returns (median block ms, all block ms, t0, t1, last result)."""
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        t0 = time.time()
        ev[0].record()
        k = 0
        for r in range(reps):
            for _ in range(n):
                last = fn(k); k += 1
            ev[r + 1].record()
        barrier()
        t1 = time.time()
        blocks = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
        if world > 1:
            t = torch.tensor(blocks, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            blocks = [float(v) for v in t.tolist()]
        return statistics.median(blocks), blocks, t0, t1, last

    # ---- step 0 self-check: forward of the INITIAL replica on batch 0 (eval mode: no dropout), compared below with the CPU arm
    #      run on the very same parameters / ids / states / popularity table ----
    model.eval()
    with torch.no_grad():
        loss_gpu_step0 = float(model(*batches[0], device).item())
    selfcheck = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.workload == "instrument":
        selfcheck = {"params": {n: p.detach().float().cpu().numpy() for n, p in model.named_parameters()},
                     "batch": {"ids": batches[0][0].view(B, 11).cpu().numpy(), "log_mask": batches[0][3].cpu().numpy(),
                               "image": batches[0][1].float().cpu().numpy(), "text": batches[0][2].float().cpu().numpy()},
                     "pop": model.pop_prob_list.detach().float().cpu().numpy()}
    model.train()

    # ---- launches per step, counted on one eager step ----
    eager = TrainStep(model, opt, use_graph=False, group=group)
    if world > 1:
        model.negatives = a.negatives
    eager(*batches[0])
    torch.cuda.synchronize()
    l0 = lib.iisan_launch_count(-1)
    eager(*batches[1])
    torch.cuda.synchronize()
    launches_per_step = lib.iisan_launch_count(-1) - l0
    dbg("eager ok")

    # ---- the timed arm: inputs resident in HBM; one captured graph per resident batch (no staging copies) ----
    keep = []                                                # graphs hold NCCL work: released explicitly at shutdown

    def measure(negatives):
        if world > 1:
            model.negatives = negatives
        if use_graph:
            runners = [TrainStep(model, opt, use_graph=True, group=group) for _ in range(n_rot)]
            keep.extend(runners)
            for r, bt in zip(runners, batches):
                r.capture(*bt)
            step = lambda i: runners[i % n_rot].replay()
        else:
            step = lambda i: eager(*batches[i % n_rot])
        for i in range(max(a.warmup, 3)):
            step(i)
        return step

    sampler = ClockSampler(local) if rank == 0 else None
    step = measure(a.negatives)
    if sampler:
        sampler.start(); time.sleep(0.3)
    dbg("warm")
    ms, blocks, t0, t1, last = timed(a.steps, step, a.reps)
    dbg("timed")
    if sampler:
        time.sleep(0.2); sampler.stop()
    clocks = sampler.summary(t0, t1) if sampler else None
    value = world * B * a.steps / (ms / 1e3)
    loss_val = float(last.item())
    other = None
    if world > 1:                                            # the other negative-pool definition, same run
        om = "local" if a.negatives == "global" else "global"
        ostep = measure(om)
        oms, oblocks, _, _, _ = timed(a.steps, ostep, a.reps)
        other = {"negatives": om, "value": world * B * a.steps / (oms / 1e3), "unit": UNIT, "ms_per_step": oms / a.steps,
                 "block_ms_min_max": [min(oblocks), max(oblocks)]}
        model.negatives = a.negatives

    # ---- per-kernel-class device time: K eager steps with CUDA events around every library launch ----
    lib.iisan_timing_enable(1)
    ms_prof, _, _, _, _ = timed(a.steps, lambda i: eager(*batches[i % n_rot]))
    lib.iisan_timing_enable(0)
    classes = {}
    for k, name in enumerate(_lib.KERNEL_CLASSES):
        tot, n = C.c_double(0), C.c_int64(0)
        lib.iisan_timing_read(k, C.byref(tot), C.byref(n))
        classes[name] = {"ms_per_step": tot.value / a.steps, "launches_per_step": n.value / a.steps}

    # ---- exposed communication (N > 1): the local-negatives step with and without its gradient all-reduce ----
    exposed = None
    if world > 1 and use_graph:
        model.negatives = "local"
        runners = [TrainStep(model, opt, use_graph=True, group=False) for _ in range(n_rot)]      # group=False: no collectives at all
        for r, bt in zip(runners, batches):
            r.capture(*bt)
        for i in range(3):
            runners[i % n_rot].replay()
        nms, _, _, _, _ = timed(a.steps, lambda i: runners[i % n_rot].replay(), max(3, a.reps // 2))
        local_ms = (ms if a.negatives == "local" else other["ms_per_step"] * a.steps) / a.steps
        # the bare collective: one fp32 all-reduce of the size of the gradient arena, back to back
        n_par = sum(p.numel() for p in model.parameters() if p.requires_grad)
        probe = torch.zeros(n_par, dtype=torch.float32, device=device)
        for _ in range(5):
            dist.all_reduce(probe, op=dist.ReduceOp.AVG)
        ar_ms, _, _, _, _ = timed(20, lambda i: dist.all_reduce(probe, op=dist.ReduceOp.AVG))
        exposed = {"ms_per_step_no_collectives": nms / a.steps, "ms_per_step_local_negatives": local_ms,
                   "bare_allreduce_ms": ar_ms / 20, "allreduce_bytes": n_par * 4,
                   "exposed_comm_ms": local_ms - nms / a.steps,
                   "note": "local in-batch negatives: captured step with the gradient all-reduce minus the same step without any collective"}
        runners = None
        model.negatives = a.negatives
    host = []
    for ids, image, text, lm in batches[:2]:
        if a.workload == "instrument":
            host.append(tuple(t.cpu().pin_memory() for t in (ids, image, text, lm)))
        else:                                                # only ids / log_mask are needed on the host (store path)
            host.append((ids.cpu().pin_memory(), image, text, lm.cpu().pin_memory()))
    e2e_steps = max(3, min(a.steps, 20))

    # ---- end to end, the product path (north_star subsystem 1): the cached states of the whole catalogue are packed once into
    #      HBM (selected layers only); a batch is (ids, log_mask) from pinned host memory, the per-item / per-layer gather runs on
    #      the device inside the captured step; the loss is read back every step ----
    e2e = None
    if not a.no_store:
        from iisan_b200.store import CachedStateStore
        plan = model.mm_encoder.plan
        # the catalogue tables are generated already packed (only the layers the towers read): the full [item_num+1, 81, 8192]
        # table of the LLaMA workload would be 30 GB
        tab = lambda n_l, d: torch.randn(item_num + 1, n_l, d, device=device, generator=gen, dtype=torch.float32).to(state_dtype)
        store = CachedStateStore(tab(len(plan.layers_img_read), W["d_img"]), tab(len(plan.layers_text_read), W["d_text"]),
                                 range(len(plan.layers_img_read)), range(len(plan.layers_text_read)), device=device, dtype=state_dtype)
        srunner = PipelinedTrainStep(model, opt, group=group, use_graph=use_graph, store=store)
        hids = [(h[0], h[3]) for h in host]
        srunner.submit(hids[0][0], log_mask=hids[0][1])

        def store_step_sync(i):
            hi, hl = hids[(i + 1) % len(hids)]
            srunner.submit(hi, log_mask=hl)                                  # H2D of the next batch's ids + log_mask
            return srunner.run().item()                                      # step + D2H read of ITS loss: the host waits for the step

        def store_step(i):
            hi, hl = hids[(i + 1) % len(hids)]
            srunner.submit(hi, log_mask=hl)                                  # H2D of the next batch's ids + log_mask
            return srunner.run_logged()                                      # step + D2H copy of its loss; returns the loss of step i - 1

        for i in range(4):
            store_step_sync(i)
        ms_sync, _, _, _, _ = timed(e2e_steps, store_step_sync, min(a.reps, 3))
        for i in range(4):
            store_step(i)
        ms_st, st_blocks, _, _, last_logged = timed(e2e_steps, store_step, min(a.reps, 5))
        e2e = {"value": world * B * e2e_steps / (ms_st / 1e3), "unit": UNIT, "ms_per_step": ms_st / e2e_steps,
               "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in hids[0]), "d2h_bytes_per_step": 4,
               "path": "HBM-resident cached-state store",
               "loss_readback": "every step's loss is copied to pinned host memory inside the timed region and read by the host one step "
                                "late (PipelinedTrainStep.run_logged): the host never waits for the step it has just launched",
               "loss_last_read": srunner.last_loss(),
               "sync_readback": {"value": world * B * e2e_steps / (ms_sync / 1e3), "ms_per_step": ms_sync / e2e_steps,
                                 "note": "same path with loss.item() after every step (the host waits for each step before it submits the next)"},
               "store_bytes_hbm": int(store.image.numel() * store.image.element_size() + store.text.numel() * store.text.element_size()),
               "note": "PipelinedTrainStep(store=CachedStateStore), the public API of the cached hidden-state path: the 7+7 selected layers "
                       "of the whole catalogue are resident in HBM; every timed step copies one batch of ids + log_mask from pinned HOST "
                       "memory; the per-item / per-layer gather of batch i+1 runs on the device on the copy stream while step i computes "
                       "(PipelinedTrainStep(prefetch_gather=True)); the captured step consumes the gathered [N, 7, 768] tensors"}
        srunner = None
        keep.append(store)

    # ---- end to end from HOST batches of the reference shapes [B,11,13,768] (Code_Cached/run.py:368-377 as is): bound by the
    #      host link (the selected layers of one batch are 121 MB in bf16) ----
    def host_batch_e2e(hb, label):
        runner = PipelinedTrainStep(model, opt, group=group, use_graph=use_graph)
        keep.append(runner)
        runner.submit(*hb[0])
        n_sel = len(set(model.mm_encoder.plan.layers_img_read)) + len(set(model.mm_encoder.plan.layers_text_read))
        pl = model.mm_encoder.plan
        nbytes = (sum(t.numel() * t.element_size() for t in (hb[0][0], hb[0][3]))
                  + B * 11 * (len(pl.layers_img_read) * W["d_img"] + len(pl.layers_text_read) * W["d_text"]) * hb[0][1].element_size())

        def fn(i):
            runner.submit(*hb[(i + 1) % len(hb)])                            # H2D of the next batch (selected layers) on the copy stream
            return runner.run().item()                                       # step on the batch submitted one call earlier + loss read-back

        for i in range(3):
            fn(i)
        ms_h, _, _, _, _ = timed(e2e_steps, fn, 3)
        return {"value": world * B * e2e_steps / (ms_h / 1e3), "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_h / e2e_steps, "host_link_gbs": nbytes / (ms_h / e2e_steps / 1e3) / 1e9, "host_dtype": label}

    e2e_host = None
    if not a.no_host_e2e and (a.workload == "instrument" or e2e is None):          # (the pinned host copies of a LLaMA-shaped batch would be 2 x 10 GB)
        if not host[0][1].is_pinned():
            host = [tuple(t.cpu().pin_memory() for t in bt) for bt in batches[:2]]
        e2e_host = host_batch_e2e(host, str(state_dtype).split(".")[-1])
        e2e_host["note"] = ("PipelinedTrainStep.submit/run with pinned HOST batch tensors of the reference shapes: every timed step issues "
                            "the H2D copy of one batch (ids, log_mask, the selected layers) and reads one loss back; the copy of batch i+1 "
                            "overlaps the step of batch i")
        if e2e is None:
            e2e = dict(e2e_host, path="host batches")
    # the same with the states as the reference's DataLoader delivers them: fp32 (dataset.py:29-34 loads the fp32 .pt files).  The
    # copy doubles (242 MB per step); on the device the selected layers are rounded to bf16 once (iisan_pack_states) and take
    # the same fused kernels.  One GPU only (2 x 0.9 GB of pinned host memory per rank).
    e2e_host_fp32 = None
    if e2e_host is not None and world == 1 and a.workload == "instrument" and state_dtype != torch.float32:
        host32 = [(h[0], h[1].float().pin_memory(), h[2].float().pin_memory(), h[3]) for h in host]
        e2e_host_fp32 = host_batch_e2e(host32, "float32")
        host32 = None

    def shutdown():
        """Release the captured graphs (they hold NCCL work) before tearing the process group down; a process that still
        cannot finalise NCCL within 20 s exits anyway (the JSON line is already out)."""
        if world == 1:
            return
        keep.clear()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        t = threading.Timer(20.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            t.cancel()

    if rank != 0:
        shutdown()
        return

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    elt = 4 if state_dtype == torch.float32 else 2
    plan = model.mm_encoder.plan
    # Dominant kernels of the hidden-state path: the fused chain forward and backward (one launch each per step).
    # Algorithmic bytes per launch: forward = every selected layer of every item read once,
    # S * (A_i*D_i + A_t*D_t) * sizeof(elt) per sample; the backward re-streams the same layers once for the gate gradients.
    alg_bytes = B * 11 * (len(plan.layers_img_read) * W["d_img"] + len(plan.layers_text_read) * W["d_text"]) * elt
    traffic, traffic_src = ncu_traffic()

    def hbm_roof(cls, kname, tkey):
        c = classes[cls]
        k_ms = c["ms_per_step"] / c["launches_per_step"] if c["launches_per_step"] > 0 else 0.0
        r = {"bound": "hbm", "kernel": kname, "achieved": alg_bytes / (k_ms / 1e3) / 1e9 if k_ms > 0 else None, "peak": hbm_peak,
             "unit": "GB/s", "traffic": traffic.get(tkey), "traffic_source": traffic_src, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": k_ms}
        r["frac"] = (r["achieved"] / hbm_peak) if r["achieved"] else None
        return r

    gen3 = classes.get("wgrad_stream", {}).get("launches_per_step", 0) > 0          # third generation: san_chain3.cu + san_lr.cu
    if classes["chain"]["launches_per_step"] > 0 and gen3:
        roof = hbm_roof("chain", "san_chain3_fwd_kernel (resident-state chain: layer-select gather + gate fusion + 3 x 7 adapters + merged "
                                 "heads, forward; writes only relu(z) [N,64] per stage and the E outputs)", "fwd")
        # The backward reads the hidden states exactly once more, inside ONE tensor-core GEMM launch (G = dz^T h for every tower,
        # stage pair and layer, plus the Gram blocks dz^T relu(z)): tensor bound.  FLOPs: 2 * N * d * 64 per (layer j, stage s >= j)
        # block, 36 blocks per intra-modal tower and 2 x 36 for the inter-modal one, + the 28 needed Gram blocks per tower.
        A_st = len(plan.stages)
        n_items = B * 11
        g_flop = 2.0 * n_items * plan.d_mm * 64 * ((A_st + 1) * (A_st + 2) / 2 - 1) * 4 + 2.0 * n_items * 64 * 64 * (A_st * (A_st + 1) / 2) * 3
        wg = classes["wgrad_stream"]
        wg_ms = wg["ms_per_step"] / wg["launches_per_step"]
        roof_bwd = {"bound": "tensor", "kernel": "umma_gemm_kernel<256, MN, MN> (the backward's one pass over the hidden states: G = dz^T h, + Gram blocks dz^T relu(z))",
                    "achieved": g_flop / (wg_ms / 1e3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s", "peak_source": peak_src,
                    "useful_flop_per_launch": g_flop, "ms_per_launch": wg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                    "hbm_gbs_of_the_algorithmic_bytes": alg_bytes / (wg_ms / 1e3) / 1e9, "traffic": traffic.get("wgrad"), "traffic_source": traffic_src,
                    "rank_space_chain_ms": classes["chain_bwd"]["ms_per_step"]}
        roof_bwd["frac"] = roof_bwd["achieved"] / tf_peak
    elif classes["chain"]["launches_per_step"] > 0:
        roof = hbm_roof("chain", "san_chain2_fwd_kernel (fused layer-select gather + gate fusion + adapter chain, forward)", "fwd")
        roof_bwd = hbm_roof("chain_bwd", "chain backward kernel (fused data/gate/bias gradients of the chain)", "bwd")
    else:
        # layered path (group layer-drop / dim alignment): the hidden states are streamed by the mix kernels, once in the forward
        # and once in the backward; all their launches of a step are summed against 2 x the algorithmic bytes
        st_ms = classes["stream"]["ms_per_step"]
        roof = {"bound": "hbm", "kernel": "mix kernels (layer-select gather + gate fusion forward, gate-gradient re-stream backward), all launches of a step",
                "achieved": 2 * alg_bytes / (st_ms / 1e3) / 1e9 if st_ms > 0 else None, "peak": hbm_peak, "unit": "GB/s", "traffic": None,
                "traffic_source": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": 2 * alg_bytes, "ms_per_launch": st_ms}
        roof["frac"] = (roof["achieved"] / hbm_peak) if roof["achieved"] else None
        roof_bwd = None
    gemm_ms = (classes["gemm"]["ms_per_step"] + classes["chain"]["ms_per_step"] + classes["chain_bwd"]["ms_per_step"] +
               classes.get("wgrad_stream", {}).get("ms_per_step", 0.0))
    # forward GEMM FLOPs of the SAN per item: adapters (down + up per active tower and stage), dim-alignment GEMMs, heads
    E = plan.emb
    f_item = 0
    for (ta, _tl, ia, _il, mi) in plan.stages:
        if ta >= 0: f_item += 4 * plan.d_text * plan.r_text
        if ia >= 0: f_item += 4 * plan.d_img * plan.r_img
        if mi >= 0:
            f_item += 4 * plan.d_mm * plan.r_mm
            if plan.n_down_project: f_item += 2 * max(plan.d_text, plan.d_img) * plan.d_mm
    ft, fi = (E, E) if plan.asym else (plan.d_text, plan.d_img)
    f_item += 2 * plan.d_text * ft + 2 * ft * E + 2 * plan.d_img * fi + 2 * fi * E + 2 * plan.d_mm * plan.d_mm + 2 * plan.d_mm * E
    roof_tensor = {"bound": "tensor", "achieved": (B * 11 * 3 * f_item / (gemm_ms / 1e3) / 1e12) if gemm_ms > 0 else None,
                   "peak": tf_peak, "unit": "TFLOP/s", "ms_per_step": gemm_ms, "flop_per_item_forward": f_item,
                   "note": "SAN adapter / alignment / head GEMM FLOPs (fwd+bwd = 3x forward) over the summed GEMM-class + chain kernel time"}
    roof_tensor["frac"] = (roof_tensor["achieved"] / tf_peak) if roof_tensor["achieved"] else None

    cpu = None
    loss_ref_step0 = None
    if world == 1 and not a.no_cpu_baseline:
        a.cpu_batch = a.cpu_batch or W["cpu_batch"]
        # a bounded sample of the same workload: at least --cpu-steps steps and ~10 s of CPU work (at most 64 steps)
        sps, sec, cores, _, n_cpu = cpu_reference_steps(a.cpu_batch, a.cpu_steps, 1, a.workload, min_seconds=a.cpu_seconds)
        cpu = {"value": sps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full train steps of B={a.cpu_batch} dense users (fp32, torch CPU, oracle port of the reference "
                         f"algorithm with vectorised negative masks -- faster than the reference's per-user Python mask loop; {sec:.2f} s/step)"}
        if selfcheck is not None:
            loss_ref_step0 = cpu_reference_loss(selfcheck["params"], selfcheck["batch"], selfcheck["pop"], item_num, a.workload)

    n_timed = a.steps * a.reps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.compute, "data": "synthetic",
        "timing": {"reps": a.reps, "statistic": "median over reps of the device time of one block of `steps` steps (max over ranks per block)",
                   "block_ms_min_max": [min(blocks), max(blocks)], "timed_region_s": t1 - t0},
        "config": {"workload": workload_name(B, str(state_dtype).split('.')[-1], item_num, a.workload), "workload_key": a.workload,
                   "negatives": ("global (all-gather)" if a.negatives == "global" else "local (reference DDP semantics)") if world > 1 else "local",
                   "l2_policy": f"inputs rotate over {n_rot} resident batches of "
                                f"{B * 11 * (W['layers_img'] * W['d_img'] + W['layers_text'] * W['d_text']) * elt / 1e6:.0f} MB (> 126 MB L2)",
                   "step_runner": "CUDA graph replay (iisan_b200.engine.TrainStep)" if use_graph else "eager",
                   "parallelism": f"dp{world}", "host_numa_binding": numa},
        "e2e": e2e, "e2e_host_batches": e2e_host, "e2e_host_batches_fp32": e2e_host_fp32,
        "gpu_launches": int(launches_per_step * n_timed),
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks,
        "roofline": (roof_tensor if a.workload == "llama_eva" else roof), "roofline_hbm": roof, "roofline_chain_bwd": roof_bwd,
        "roofline_tensor": roof_tensor,
        "kernel_classes": classes, "ms_per_step_eager_with_kernel_events": ms_prof / a.steps,
        "cpu_baseline": cpu, "loss": loss_val,
        "loss_gpu_step0": loss_gpu_step0, "loss_ref_step0": loss_ref_step0,
        # generation of the fused side-adapter kernels that ran: 3 = resident-state forward + low-rank adjoint backward,
        # 2 = the stash-based chain (data-parallel steps use it until the 8-GPU fault of generation 3 is understood: DESIGN 4.8)
        "chain_generation": int(lib.iisan_debug_chain_generation(0)),
    }
    if other is not None:
        line["other_negatives"] = other
    if exposed is not None:
        line["exposed_comm"] = exposed
    print(json.dumps(line), flush=True)
    shutdown()


def supervised(a):
    """Single-GPU arm: the measurement runs in a child process.  Generation 3 of the fused side-adapter path has an intermittent,
    not yet explained device fault (about one bench process in sixteen on one B200: DESIGN.md 4.8, profiles/r02_8gpu_gen3_fault.md);
    a CUDA fault is not recoverable inside a process, so a faulted child is re-run ONCE with generation 2 and the line says so."""
    import subprocess
    base = [sys.executable, os.path.abspath(__file__), *sys.argv[1:], "--child"]
    notes = []
    for attempt, gen in enumerate((os.environ.get("IISAN_B200_CHAIN_GEN", "3"), "2")):
        env = dict(os.environ, IISAN_B200_CHAIN_GEN=gen)
        try:      # a child that hangs is killed (subprocess.run does that on timeout) and counts as failed
            r = subprocess.run(base, env=env, stdout=subprocess.PIPE, text=True, timeout=float(os.environ.get("IISAN_BENCH_CHILD_TIMEOUT", "1500")))
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            r = subprocess.CompletedProcess(base, returncode=-9, stdout=out)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode == 0 and lines:
            line = json.loads(lines[-1])
            line["chain_generation"] = int(gen)
            if notes:
                line["fallback"] = notes
            print(json.dumps(line), flush=True)
            return 0
        notes.append({"chain_generation": int(gen), "returncode": r.returncode, "note": "child process failed; re-run with generation 2"})
    print(json.dumps({"metric": METRIC, "error": "both attempts failed", "attempts": notes}), flush=True)
    return 1


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    elif int(os.environ.get("WORLD_SIZE", "1")) == 1 and not a.child:
        sys.exit(supervised(a))
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
