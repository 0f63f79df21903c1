#!/usr/bin/env python
"""Size-matched streaming ceilings on this GPU (development aid): what a plain fill / copy / cast
kernel of the SAME byte volume as one bench step achieves, timed like tools/kbench.py.  Short
kernels pay a fixed launch / ramp / drain cost that the 4 GB copy behind MEASURED_PEAKS.json does not."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200  # noqa: E402

dev = torch.device("cuda:0")


def t(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        torch.cuda._sleep(400000)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) * 1e3 for a, b in ev)


for (W, H, B) in [(640, 360, 64), (1280, 720, 32)]:
    n = B * H * W
    sets = 4
    outs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(sets)]
    gts = [torch.zeros(n, dtype=torch.int64, device=dev) for _ in range(sets)]
    srcs = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(sets)]
    i = [0]

    def nxt():
        i[0] = (i[0] + 1) % sets
        return i[0]
    us = t(lambda: outs[nxt()].fill_(0.25))
    print(f"{W}x{H} B{B} fill fp32 ({n*4/1e6:.0f} MB write)          {us:7.1f} us {n*4/us/1e3:6.0f} GB/s")
    us = t(lambda: outs[nxt()].copy_(srcs[i[0]]))
    print(f"{W}x{H} B{B} copy fp32->fp32 ({n*8/1e6:.0f} MB r+w)       {us:7.1f} us {n*8/us/1e3:6.0f} GB/s")
    us = t(lambda: outs[nxt()].copy_(gts[i[0]]))
    print(f"{W}x{H} B{B} cast int64->fp32 ({n*12/1e6:.0f} MB r+w) = C2 bytes {us:7.1f} us {n*12/us/1e3:6.0f} GB/s")
    lib = sfh_b200._lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    for ctas in (148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32):
        us = t(lambda: lib.sfh_debug_stream_cast(gts[nxt()].data_ptr(), outs[i[0]].data_ptr(), n, ctas, st))
        print(f"{W}x{H} B{B} own int64->fp32 stream kernel, {ctas:5d} CTAs ({n*12/1e6:.0f} MB)   {us:7.1f} us {n*12/us/1e3:6.0f} GB/s")
    us = t(lambda: gts[nxt()].sum())
    print(f"{W}x{H} B{B} sum int64 ({n*8/1e6:.0f} MB read)           {us:7.1f} us {n*8/us/1e3:6.0f} GB/s")
