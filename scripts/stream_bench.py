"""HBM-bound DIIS steps at BASELINE sizes: fused multi-dot (PulayDiisScf.hh:440-463) and m-way extrapolation
(PulayDiisScf.hh:395-433).  Prints achieved GB/s against the algorithmic bytes (k+1) * 8 n^2 of each."""
import ctypes as C
import sys

sys.path.insert(0, '.')
import numpy as np
import torch

from gscf_b200 import lib as L

for n, count in [(4096, 10), (1024, 8)]:
    ld = L.lib.gscf_b200_padded_ld(n)
    nslots = count + 1
    ring = torch.randn(nslots * n * ld, dtype=torch.float64, device="cuda")
    enew = torch.randn(n * ld, dtype=torch.float64, device="cuda")
    slots = np.arange(count, dtype=np.int32)
    out = torch.zeros(16, dtype=torch.float64, device="cuda")
    wsz = L.lib.gscf_b200_frob_multidot_worksize(count, n, n, 1)
    work = torch.zeros(max(1, wsz // 8), dtype=torch.float64, device="cuda")
    coef = torch.randn(16, dtype=torch.float64, device="cuda")
    res = torch.zeros(n * ld, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    sp = slots.ctypes.data_as(C.POINTER(C.c_int))

    def multidot():
        L.check(L.lib.gscf_b200_frob_multidot(enew.data_ptr(), ring.data_ptr(), n * ld, sp, count, n, n, ld, n * ld,
                                              nslots * n * ld, 1, out.data_ptr(), 16, work.data_ptr(), wsz, None))

    def axpby():
        L.check(L.lib.gscf_b200_axpby_n(res.data_ptr(), n * ld, ring.data_ptr(), n * ld, nslots * n * ld, sp, count,
                                        coef.data_ptr(), 16, n, n, ld, 1, None))

    for name, fn, nbytes in [("multidot", multidot, (count + 1) * 8.0 * n * n), ("axpby_n", axpby, (count + 1) * 8.0 * n * n)]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(6):
            flush.zero_()
            torch.cuda.synchronize()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[1:]))
        print(f"{name:9s} n={n} k={count}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic {(nbytes / 1e6):.0f} MB, L2 flushed)")
