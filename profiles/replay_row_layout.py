#!/usr/bin/env python
"""Replay ring layout experiment: five SoA arrays (the reference's layout, buffer.py:7-11, and
what libboatenv ships) against ONE packed 128-byte row per transition
    [state 11 f32 | action | reward | new_state 11 f32 | done as f32 | 7 f32 padding]  = 32 floats.

The packed ring is emulated with plain torch ops on a [N, 32] float tensor (store = one contiguous row-block copy,
gather = index_select), which is how a packed implementation would move its bytes; the SoA numbers come from the
product kernels (boatreplay_store / boatreplay_sample).  Prints one JSON line per case: ms, rows/s, and the DRAM bytes
per row each layout has to move at minimum.  Run on the GPU box:  python profiles/replay_row_layout.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sac_agent_b200 as S  # noqa: E402


def timed(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    torch.cuda.set_device(0)
    cap, n_store = 16 << 20, 4 << 20
    s = torch.randn(n_store, 11, device="cuda")
    a = torch.randn(n_store, 1, device="cuda")
    r = torch.randn(n_store, device="cuda")
    d = torch.zeros(n_store, dtype=torch.uint8, device="cuda")
    buf = S.ReplayBuffer(cap, (11,), 1, precision="fp32", device=0, as_torch=True)
    for _ in range(4):
        buf.store_batch(s, a, r, s, d)           # fill the ring: gathers below hit all of it
    packed_ring = torch.randn(cap, 32, device="cuda")
    rows = torch.randn(n_store, 32, device="cuda")
    out = []
    ms = timed(lambda: buf.store_batch(s, a, r, s, d))
    out.append({"case": "store_4M_rows_soa (product kernel)", "ms": ms, "rows_per_s": n_store / (ms * 1e-3),
                "bytes_written_per_row": 97})
    ms = timed(lambda: packed_ring[:n_store].copy_(rows))
    out.append({"case": "store_4M_rows_packed_128B (torch copy of a [4M, 32] block)", "ms": ms,
                "rows_per_s": n_store / (ms * 1e-3), "bytes_written_per_row": 128})
    for batch in (1024, 1 << 20):
        outs = buf.sample_buffer(batch)
        outs = (outs[0], outs[1], outs[2], outs[3], outs[4].to(torch.uint8))
        ms = timed(lambda: buf.sample_buffer(batch, out=outs), iters=50)
        out.append({"case": f"gather_{batch}_soa (product kernel, Philox indices in-kernel)", "ms": ms,
                    "rows_per_s": batch / (ms * 1e-3), "useful_bytes_per_row": 97,
                    "dram_bytes_per_row_min": "5 arrays x 64-byte sector pairs for 44 / 44 / 4 / 4 / 1 useful bytes"})
        idx = torch.randint(0, cap, (batch,), device="cuda")
        dst = torch.empty(batch, 32, device="cuda")
        ms = timed(lambda: torch.index_select(packed_ring, 0, idx, out=dst), iters=50)
        out.append({"case": f"gather_{batch}_packed_128B (torch.index_select of [16M, 32] rows)", "ms": ms,
                    "rows_per_s": batch / (ms * 1e-3), "useful_bytes_per_row": 97, "dram_bytes_per_row_min": 128})
    for line in out:
        print(json.dumps(line), flush=True)
    buf.close()


if __name__ == "__main__":
    main()
