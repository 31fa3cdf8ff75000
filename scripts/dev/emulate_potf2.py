"""Lane-level numpy emulation of potf2_panels (miniframe_b200/csrc/mf_potrf.cu): checks the index
algebra (fragment layout, shuffle sources, k-permuted DMMA operands, block layout of Ld, Dinv)
against numpy's Cholesky before any GPU time is spent."""
import numpy as np

AW_LD = 65


def ldblk(bi, bj):
    return (bi * (bi + 1) // 2 + bj) * 64


def dmma(d, a, b):
    """d: (32, 2) C/D fragment, a, b: (32,) A/B fragments.  lane 4g+t holds A[g][t], B[t][g], D[g][2t..]."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t] = a[lane]; B[t, g] = b[lane]
    P = A @ B
    out = d.copy()
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane, 0] += P[g, 2 * t]; out[lane, 1] += P[g, 2 * t + 1]
    return out


def potf2(Aw, nv):
    Ld = np.full(36 * 64, np.nan); Dinv = np.full(8 * 64, np.nan); rsbuf = np.zeros(64)
    lanes = np.arange(32); G = lanes >> 2; T = lanes & 3

    def ld2(block):         # double2 per lane at block + g*8 + 2t
        return np.stack([Ld[block + G * 8 + 2 * T], Ld[block + G * 8 + 2 * T + 1]], 1)

    for warp in range(8):    # panels in order (the mbarrier chain serialises them)
        c0 = 8 * warp + 2 * T

        def load_neg(bi):
            r = 8 * bi + G
            v0 = Aw[r, c0].copy(); v1 = Aw[r, c0 + 1].copy()
            return np.stack([-v0, -v1], 1)

        live0, live1 = c0 < nv, c0 + 1 < nv

        dg = load_neg(warp)
        for q in range(warp):
            bq = ld2(ldblk(warp, q))
            dg = dmma(dg, bq[:, 0], bq[:, 0]); dg = dmma(dg, bq[:, 1], bq[:, 1])
        dg = -dg
        dg[:, 0] = np.where(live0, dg[:, 0], (8 * warp + G == c0) * 1.0)
        dg[:, 1] = np.where(live1, dg[:, 1], (8 * warp + G == c0 + 1) * 1.0)
        for c in range(8):
            col = dg[:, c & 1]
            d = col[4 * c + (c >> 1)]
            assert d > 0
            rs = 1.0 / np.sqrt(d); rsbuf[8 * warp + c] = rs
            sel = T == (c >> 1)
            newcol = np.where(G > c, col * rs, np.where(G == c, d * rs, 0.0))
            col = np.where(sel, newcol, col); dg[:, c & 1] = col
            if c < 7:
                l0 = col[8 * T + (c >> 1)]; l1 = col[8 * T + 4 + (c >> 1)]; lr = col[4 * G + (c >> 1)]
                dg[:, 0] = np.where((2 * T > c) & live0, dg[:, 0] - lr * l0, dg[:, 0])
                dg[:, 1] = np.where((2 * T + 1 > c) & live1, dg[:, 1] - lr * l1, dg[:, 1])
        dg[:, 0] = np.where(2 * T > G, 0.0, dg[:, 0]); dg[:, 1] = np.where(2 * T + 1 > G, 0.0, dg[:, 1])
        blk = ldblk(warp, warp)
        Ld[blk + G * 8 + 2 * T] = dg[:, 0]; Ld[blk + G * 8 + 2 * T + 1] = dg[:, 1]
        for lane in range(8):
            x = np.zeros(8)
            for i in range(8):
                s = -1.0 if i == lane else 0.0
                for k in range(i):
                    s -= Ld[blk + i * 8 + k] * x[k]
                x[i] = s * rsbuf[8 * warp + i]
            for i in range(8):
                Dinv[warp * 64 + i * 8 + lane] = x[i]
        dv = np.stack([Dinv[warp * 64 + G * 8 + 2 * T], Dinv[warp * 64 + G * 8 + 2 * T + 1]], 1)
        for bi in range(warp + 1, 8):
            a = load_neg(bi)
            for q in range(warp):
                bq = ld2(ldblk(warp, q)); aq = ld2(ldblk(bi, q))
                a = dmma(a, aq[:, 0], bq[:, 0]); a = dmma(a, aq[:, 1], bq[:, 1])
            a[:, 0] = np.where(live0, a[:, 0], 0.0); a[:, 1] = np.where(live1, a[:, 1], 0.0)
            x = np.zeros((32, 2))
            x = dmma(x, a[:, 0], dv[:, 0]); x = dmma(x, a[:, 1], dv[:, 1])
            b2 = ldblk(bi, warp)
            Ld[b2 + G * 8 + 2 * T] = x[:, 0]; Ld[b2 + G * 8 + 2 * T + 1] = x[:, 1]
    return Ld, Dinv


def unpack(Ld):
    L = np.zeros((64, 64))
    for bi in range(8):
        for bj in range(bi + 1):
            L[8 * bi:8 * bi + 8, 8 * bj:8 * bj + 8] = Ld[ldblk(bi, bj):ldblk(bi, bj) + 64].reshape(8, 8)
    return L


rng = np.random.default_rng(0)
for nv in (64, 28, 8, 1, 63):
    X = rng.standard_normal((nv, nv + 5)); K = X @ X.T + 0.1 * np.eye(nv)
    r = rng.standard_normal(nv)
    Aw = rng.standard_normal((128, AW_LD)) * 7.0           # garbage everywhere the kernel may not rely on
    Aw[:nv, :nv] = np.tril(K) + np.triu(rng.standard_normal((nv, nv)), 1)   # upper triangle: garbage
    if nv < 64:
        Aw[nv, :nv] = r                                    # the residual row
        Aw[nv + 1:, :] = 0.0                               # rows beyond it are zero in the kernel
    Ld, Dinv = potf2(Aw, nv)
    L = unpack(Ld)
    Lref = np.linalg.cholesky(K)
    err = np.abs(L[:nv, :nv] - Lref).max()
    e2 = 0.0
    if nv < 64:
        z = np.linalg.solve(Lref, r)
        e2 = np.abs(L[nv, :nv] - z).max()
        assert np.allclose(L[nv:, nv:], np.eye(64 - nv)), "identity tail"
    Lfull = L.copy()
    e3 = 0.0
    for b in range(8):
        blk = Lfull[8 * b:8 * b + 8, 8 * b:8 * b + 8]
        e3 = max(e3, np.abs(Dinv[b * 64:(b + 1) * 64].reshape(8, 8) + np.linalg.inv(blk)).max())
    print("nv %2d: |L - chol| %.2e  |z| %.2e  |Dinv + inv| %.2e" % (nv, err, e2, e3))
    assert err < 1e-10 and e2 < 1e-9 and e3 < 1e-9
print("emulation ok")
