import sys
sys.path.insert(0, '.')
import torch
from gscf_b200 import lib as L
for (m, n, k, ta, tb) in [(4096, 4096, 4096, 0, 0), (4096, 4096, 4096, 0, 1), (4096, 4096, 4096, 1, 0), (4096, 4096, 512, 0, 1),
                          (4096, 4096, 64, 0, 1), (1024, 1024, 1024, 0, 0), (1024, 1024, 128, 0, 1), (2048, 2048, 2048, 0, 0),
                          (8192, 8192, 2048, 0, 0), (64, 4096, 4096, 1, 0), (4096, 4096, 64, 0, 0), (2048, 4096, 64, 0, 0), (4096, 512, 4096, 0, 0), (4096, 4096, 512, 0, 0)]:
    A = torch.randn(m * k, dtype=torch.float64, device="cuda")
    B = torch.randn(k * n, dtype=torch.float64, device="cuda")
    Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    lda = k if ta else m
    ldb = n if tb else k
    def run():
        L.check(L.lib.gscf_b200_dgemm(ta, tb, m, n, k, 1.0, A.data_ptr(), lda, 0, B.data_ptr(), ldb, 0, 0.0, Cm.data_ptr(), m, 0, 1, None))
    for _ in range(2): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    reps = 5
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"dgemm m={m} n={n} k={k} ta={ta} tb={tb}: {ms:.3f} ms  {2.0*m*n*k/ms/1e9:.2f} TFLOP/s")
