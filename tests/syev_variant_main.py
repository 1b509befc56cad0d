"""Child process of test_syev_fused_variants: the eigensolver's path selection is read from the
environment once per process, so each variant (forced global / blocked / clustered / no solo) gets its own."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gscf_b200 import lib as L  # noqa: E402


def main():
    n, batch = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(n * 7 + batch)
    ld = L.lib.gscf_b200_padded_ld(n)
    mats = []
    for b in range(min(batch, 3)):
        a = rng.standard_normal((n, n))
        a = a + a.T
        if b == 1:
            q, _ = np.linalg.qr(rng.standard_normal((n, n)))
            a = (q * np.repeat(np.arange(n // 4 + 1), 4)[:n].astype(float)) @ q.T
            a = 0.5 * (a + a.T)
        mats.append(a)
    buf = np.zeros((batch, n, ld))
    for b in range(batch):
        buf[b, :, :n] = mats[b % len(mats)]
    dA = torch.from_numpy(buf.ravel().copy()).cuda()
    dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
    dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
    info = torch.full((batch,), 7, dtype=torch.int32, device="cuda")
    wsz = L.lib.gscf_b200_syev_worksize(n, batch, L.EIG_TRIDIAG_DC)
    work = torch.zeros(max(1, wsz // 8 + 1), dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_syev(L.EIG_TRIDIAG_DC, n, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld,
                                 ld * n, batch, work.data_ptr(), wsz, info.data_ptr(), None))
    torch.cuda.synchronize()
    assert info.cpu().numpy().tolist() == [0] * batch, info
    w = dw.cpu().numpy().reshape(batch, n)
    Z = dZ.cpu().numpy().reshape(batch, n, ld)
    for b in sorted({0, min(1, batch - 1), min(2, batch - 1), batch - 1}):
        a = mats[b % len(mats)]
        wref = np.linalg.eigvalsh(a)
        scale = max(1.0, np.abs(wref).max())
        z = Z[b][:, :n].T
        assert np.abs(w[b] - wref).max() < 1e-13 * scale * n, ("eigenvalues", b)
        assert np.abs(z.T @ z - np.eye(n)).max() < 1e-13 * n, ("orthogonality", b)
        assert np.abs(a @ z - z * w[b]).max() < 1e-13 * scale * n, ("residual", b)
    print("ok")


if __name__ == "__main__":
    main()
