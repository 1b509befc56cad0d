import sys
sys.path.insert(0, '.')
import numpy as np, torch
from gscf_b200 import lib as L
for n, batch in [(300, 1), (100, 3), (1024, 1)]:
    ld = L.lib.gscf_b200_padded_ld(n)
    rng = np.random.default_rng(1)
    a = rng.standard_normal((n, n)); a = a + a.T
    buf = np.zeros((batch, n, ld)); buf[:, :, :n] = a
    buf[0, 5, 7] = np.nan; buf[0, 7, 5] = np.nan
    dA = torch.from_numpy(buf.ravel().copy()).cuda()
    dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
    dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    wsz = L.lib.gscf_b200_syev_worksize(n, batch, 2)
    work = torch.zeros(max(1, wsz // 8 + 1), dtype=torch.float64, device="cuda")
    st = L.lib.gscf_b200_syev(2, n, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld, ld * n, batch, work.data_ptr(), wsz, info.data_ptr(), None)
    torch.cuda.synchronize()
    print(n, batch, "status", st, "info", info.cpu().numpy(), "finite w0:", bool(np.isfinite(dw.cpu().numpy()[:n]).all()))
