"""Round-2 measurement #1: what bandwidth do the hidden-state tile loads of the fused chain kernels reach on their
own?  iisan_probe_tile_stream issues exactly those TMA box loads (128 rows x 128 B at the row pitch of the [N, 13, 768] bf16
states) through a ring and nothing else; the same number of tiles is then read from a tile-contiguous layout.

    gpurun -- python scripts/probe_tile_stream.py [B]          # JSON lines: GB/s per (layout, ring slots, passes)

NOT YET RUN on a GPU (written after the round-1 GPU budget was spent).  Expected reading: if `strided` at 5 slots is far below
`contiguous` and below ~60 % of the measured copy bandwidth (MEASURED_PEAKS.json), in-place streaming of the reference layout is
granularity-bound and the HBM-resident store should keep tile-contiguous tables; if both are high, the chain kernels' time is
elsewhere (per-chunk hand-over chain).
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    from iisan_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    layers, d = 13, 768
    sel = [0, 2, 4, 6, 8, 10, 12]
    n_rows = 2 * B * 11                                        # image rows, then text rows: the two modalities of one batch
    n_bufs = 3                                                 # rotate > L2
    bufs = [torch.randn(n_rows, layers, d, device=dev).bfloat16() for _ in range(n_bufs)]
    tiles = (n_rows + 127) // 128
    n_tiles = tiles * len(sel) * (d // 64)
    cont = [torch.randn(n_tiles * 128, 64, device=dev).bfloat16() for _ in range(n_bufs)]
    sink = torch.zeros(8, dtype=torch.uint8, device=dev)
    arr = (C.c_int32 * len(sel))(*sel)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(buf, slots, repeat, contiguous):
        _lib.check(lib.iisan_probe_tile_stream(C.c_void_p(buf.data_ptr()), n_rows, layers, d, arr, len(sel), slots, repeat,
                                               int(contiguous), C.c_void_p(sink.data_ptr()), st), "iisan_probe_tile_stream")

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    for contiguous in (False, True):
        for slots in (2, 3, 5, 8, 12):
            for repeat in (1, 2):
                src = cont if contiguous else bufs
                for i in range(3):
                    run(src[i % n_bufs], slots, repeat, contiguous)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 20
                e0.record()
                for i in range(n):
                    run(src[i % n_bufs], slots, repeat, contiguous)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                unique = n_tiles * 16384                        # bytes that must come from HBM per launch
                print(json.dumps({"layout": "contiguous" if contiguous else "strided", "slots": slots, "passes": repeat, "ctas": tiles,
                                  "ms": ms, "hbm_gbs": unique / ms / 1e6, "tma_gbs": unique * repeat / ms / 1e6,
                                  "peaks": peaks.get("hbm_gbs") or peaks}), flush=True)


if __name__ == "__main__":
    main()
