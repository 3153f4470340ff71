"""Where do the fused chain kernels wait?  (DESIGN.md 4.1, round-2 measurement #2.)  Needs the instrumented build variant:

    IISAN_B200_BUILD_VARIANT=trace python -m iisan_b200.build                 # -> iisan_b200/lib/libiisan_b200_trace.so
    gpurun -- 'IISAN_B200_LIB=$PWD/iisan_b200/lib/libiisan_b200_trace.so python scripts/chain_trace.py [B]'

The middle CTA of every tower accumulates, per role (weight producer / MMA + store thread / data producer / one epilogue thread)
and wait site, the clock64 cycles spent waiting and the number of waits (san_chain.cu, CH_T0 / CH_T1).  Printed per 64-column
chunk step: cycles waited at each site, and the lifetime of the role's thread divided by the number of chunk steps.
NOT YET RUN on a GPU (written after the round-1 GPU budget was spent); the default build is byte-identical without the flag.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch import nn

ROLES = ("weight_producer", "mma_store_thread", "data_producer", "epilogue_thread")
SITES = {
    0: {0: "w_empty"},
    1: {0: "xk_full (epilogue output)", 1: "bulk wait_group.read 0", 2: "bulk wait_group 3", 3: "2 x tcgen05.commit", 4: "xk_full (final stage)",
        5: "w_full", 6: "z_ready", 7: "u_empty", 15: "lifetime"},
    2: {0: "h_empty"},
    3: {0: "xk_empty", 1: "h_full (ring tile)", 2: "z_full", 3: "u_full", 15: "lifetime"},
}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    path = os.environ.get("IISAN_B200_LIB")
    if not path or "trace" not in os.path.basename(path):
        raise SystemExit("set IISAN_B200_LIB to the trace build (see the docstring)")
    from iisan_b200 import _lib, model as pkg
    from iisan_b200.config import default_args
    from iisan_b200.precision import set_compute_mode
    _lib.load()
    raw = C.CDLL(path)
    raw.iisan_debug_chain_trace_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    args = default_args()

    class ImgStub(nn.Module):
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(768, args.embedding_dim)

    m = pkg.ModelMM(args, 1000, True, ImgStub(), nn.Identity(), [1.0] * 1001)
    m.mm_encoder = pkg.IISANAdaptedMModel(m.mm_encoder, args)
    m = m.to(dev)
    set_compute_mode("bf16")
    san = m.mm_encoder
    gen = torch.Generator(device=dev).manual_seed(2)
    batches = [(torch.randn(B * 11, 13, 768, device=dev, generator=gen).bfloat16(),
                torch.randn(B * 11, 13, 768, device=dev, generator=gen).bfloat16()) for _ in range(3)]

    def step(i):
        m.zero_grad(set_to_none=True)
        out = san.embed(*batches[i % 3])
        out.sum().backward()

    n_words = 2 * 3 * 4 * 16 * 2
    buf = (C.c_ulonglong * n_words)()
    for i in range(3):
        step(i)
    _lib.check(raw.iisan_debug_chain_trace_read(buf, 1), "trace reset")
    n = 10
    for i in range(n):
        step(i)
    _lib.check(raw.iisan_debug_chain_trace_read(buf, 0), "trace read")
    A, NC = 7, 12
    steps = (A + 1) * NC                                      # chunk steps per launch and CTA
    out = {}
    for p, pname in enumerate(("forward", "backward")):
        for t, tname in enumerate(("text", "image", "inter-modal")):
            for r, rname in enumerate(ROLES):
                for s in range(16):
                    base = ((((p * 3 + t) * 4 + r) * 16) + s) * 2
                    cyc, cnt = buf[base], buf[base + 1]
                    if cnt == 0:
                        continue
                    out.setdefault(pname, {}).setdefault(tname, {}).setdefault(rname, {})[SITES.get(r, {}).get(s, f"site{s}")] = {
                        "cycles_per_chunk_step": cyc / n / steps, "waits_per_launch": cnt / n}
    print(json.dumps({"B": B, "launches": n, "chunk_steps_per_launch": steps, "trace": out}, indent=1))


if __name__ == "__main__":
    main()
