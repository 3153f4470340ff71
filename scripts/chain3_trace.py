"""Where does the third-generation chain forward spend its time?  Needs the trace build:
    IISAN_B200_BUILD_VARIANT=trace python -m iisan_b200.build
    gpurun -- 'IISAN_B200_LIB=$PWD/iisan_b200/lib/libiisan_b200_trace.so python scripts/chain3_trace.py [B]'
Cycles per chunk step (96 per launch) at every lap site of every role, middle CTA of each tower, last launch."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ROLES = ["weight_producer", "mma_down", "h_producer", "mma_up", "epi_par0_half0", "epi_par0_half1", "epi_par1_half0", "epi_par1_half1"]
SITES = {0: {0: "w_empty wait"}, 1: {0: "w_full wait", 1: "x_full wait", 2: "issue + commits"}, 2: {0: "d_empty wait"},
         3: {0: "w_full wait", 1: "u_empty wait", 2: "issue + commits", 3: "z_ready wait"}}
EPI = {0: "u_full wait", 1: "tcgen05.ld U + x, release U", 2: "bias add", 3: "h tile wait", 4: "h mix + pack + release", 5: "x write + publish",
       6: "z phase (wait + epilogue)"}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    path = os.environ.get("IISAN_B200_LIB")
    if not path or "trace" not in os.path.basename(path):
        raise SystemExit("set IISAN_B200_LIB to the trace build")
    import bench
    from iisan_b200 import _lib
    _lib.load()
    raw = C.CDLL(path)
    dev = torch.device("cuda", 0)
    model, _, _ = bench.build_model(dev, "bf16")
    san = model.mm_encoder.eval()
    N = B * 11
    g = torch.Generator(device=dev).manual_seed(5)
    batches = [(torch.randn(N, 13, 768, device=dev, generator=g).bfloat16(), torch.randn(N, 13, 768, device=dev, generator=g).bfloat16())
               for _ in range(3)]
    with torch.no_grad():
        for i in range(5):
            san.embed(*batches[i % 3])
    buf = (C.c_uint * (3 * 8 * 8))()
    _lib.check(raw.iisan_debug_chain3_trace_read(buf), "trace read")
    out = {}
    for t, tname in enumerate(("text", "image", "inter-modal")):
        for r, rname in enumerate(ROLES):
            names = SITES.get(r, EPI)
            d = {}
            for s in range(8):
                v = buf[(t * 8 + r) * 8 + s]
                if v:
                    # producers / MMA threads: per chunk step (96 per launch); epilogue groups: per chunk of their parity (48 per launch)
                    div = 48 if r >= 4 else 96
                    d["lifetime" if s == 7 else names.get(s, f"site{s}")] = round(v / div, 1)
            out.setdefault(tname, {})[rname] = d
    print(json.dumps({"B": B, "cycles_per_chunk": out}, indent=1))


if __name__ == "__main__":
    main()
