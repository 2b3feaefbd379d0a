"""Measured int8 tensor-core peak of this B200 through the library path (torch._int_mm -> cuBLASLt), the denominator
of the Ozaki-scheme go/no-go estimate in DESIGN.md (FP64 emulation of the coefficient contraction on tcgen05 int8).
usage: python profiles/microbench/int8_peak.py"""
import torch
dev = torch.device("cuda:0")
for n in (4096, 8192, 16384):
    a = torch.randint(-127, 127, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-127, 127, (n, n), dtype=torch.int8, device=dev)
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"int8 {n}^3: {ms:.3f} ms  {2 * n ** 3 / ms / 1e9:.1f} TOP/s")
# the contraction's own shape through the library: (1664 x 192) . (192 x N) -> int32, output-bound
m, k, npts = 1664, 192, 1 << 20
a = torch.randint(-127, 127, (m, k), dtype=torch.int8, device=dev)
b = torch.randint(-127, 127, (k, npts), dtype=torch.int8, device=dev)
for _ in range(2):
    torch._int_mm(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    torch._int_mm(a, b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"int8 ({m} x {k}) . ({k} x 2^20) one slice pair: {ms:.3f} ms  {2 * m * k * npts / ms / 1e9:.1f} TOP/s, "
      f"int32 output {m * npts * 4 / ms / 1e6:.0f} GB/s  (x 21 slice pairs for 40-bit accuracy if not fused)")
