"""Run standalone eigensolves through the C ABI (profiling / correctness target).

usage: eig_only.py n [reps] [method] [batch] [check]
"""
import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
from gscf_b200 import lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
method = int(sys.argv[3]) if len(sys.argv) > 3 else 2
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
check = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rng = np.random.default_rng(0)
ld = L.lib.gscf_b200_padded_ld(n)
buf = np.zeros((batch, n, ld))
mats = []
for b in range(batch):
    if b < 3:
        a = rng.standard_normal((n, n)); a = a + a.T
        if b == 1:  # clustered spectrum / deflation-heavy
            q, _ = np.linalg.qr(rng.standard_normal((n, n)))
            a = (q * np.repeat(np.arange(n // 8 + 1), 8)[:n]) @ q.T
            a = 0.5 * (a + a.T)
        mats.append(a)
    else:
        a = mats[b % 3]
    buf[b, :, :n] = a
dA0 = torch.from_numpy(buf.ravel().copy()).cuda()
dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
info = torch.zeros(batch, dtype=torch.int32, device="cuda")
wsz = L.lib.gscf_b200_syev_worksize(n, batch, method)
work = torch.zeros(max(1, wsz // 8 + 1), dtype=torch.float64, device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    dA = dA0.clone()
    torch.cuda.synchronize()
    ev0.record()
    L.check(L.lib.gscf_b200_syev(method, n, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld, ld * n, batch,
                                 work.data_ptr(), wsz, info.data_ptr(), None))
    ev1.record()
    torch.cuda.synchronize()
    print(f"n={n} batch={batch} method={method} rep={r} {ev0.elapsed_time(ev1):.3f} ms", flush=True)
print("info", info.cpu().numpy()[:8])
if check:
    W = dw.cpu().numpy().reshape(batch, n)
    Z = dZ.cpu().numpy().reshape(batch, n, ld)[:, :, :n].transpose(0, 2, 1)  # Z[b][:, k] eigenvector k
    for b in sorted(set([0, min(1, batch - 1), min(2, batch - 1), batch - 1])):
        a = mats[b % 3] if b >= 3 else mats[b]
        w, z = W[b], Z[b]
        nrm = np.abs(a).max() * n
        res = np.abs(a @ z - z * w).max() / nrm
        orth = np.abs(z.T @ z - np.eye(n)).max()
        werr = np.abs(w - np.linalg.eigvalsh(a)).max() / np.abs(a).max()
        print(f"  mat {b}: eig err {werr:.2e}  residual {res:.2e}  orth {orth:.2e}")
