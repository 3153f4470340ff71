"""Where does the second-generation chain BACKWARD wait?  (trace build, see scripts/chain2_trace.py)"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ROLES = ["weight_producer", "mma_thread", "data_producer", "store_warp", "epi_group0", "epi_group1", "epi_group2", "epi_group3"]
SITES = {0: {0: "w_empty"}, 1: {0: "w_full", 1: "x_full", 2: "z_ready", 3: "u_empty", 4: "commits", 5: "cs_empty"},
         2: {0: "d_empty", 1: "stash complete"}, 3: {0: "x_full", 1: "wait_group.read", 2: "wait_group 2", 3: "cs_full"}}
EPI = {0: "d_full", 1: "x_empty", 2: "z_full", 3: "u_full", 4: "tcgen05.st+wait", 5: "fence.proxy.async", 6: "tcgen05.ld+wait"}


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
    for i in range(5):
        san.zero_grad(set_to_none=True)
        san.embed(*batches[i % 3]).sum().backward()
    buf = (C.c_uint * (3 * 8 * 8))()
    _lib.check(raw.iisan_debug_chain2_bwd_trace_read(buf), "trace read")
    steps = 8 * 12
    out = {}
    for t, tname in enumerate(("text", "image", "inter-modal")):
        for r, rname in enumerate(ROLES):
            names = SITES.get(r, EPI)
            d = {}
            for s in range(8):
                v = buf[(t * 8 + r) * 8 + s]
                if v:
                    d["lifetime" if s == 7 else names.get(s, f"site{s}")] = round(v / steps, 1)
            out.setdefault(tname, {})[rname] = d
    print(json.dumps({"B": B, "cycles_per_chunk_step": out}, indent=1))


if __name__ == "__main__":
    main()
