#!/usr/bin/env python
"""CPU emulation of k_tile_z's data flow, thread by thread, with the index arithmetic transcribed
from csrc/srb_kernels_tile.cuh (buffers A / B poisoned with NaN between phases; the in-place
horizontal pass run with the threads in REVERSED order after the tail pre-load, to show that the
barrier placement is sufficient): the gradient of interior tiles must match the oracle's data term.
This is how the kernel's design was checked before its first (and, in round 1, only) GPU run.
    python tools/emulate_tile_z.py [KH]      # PSF half width 1..4, default 3"""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sr_oracle as o

KH = int(sys.argv[1]) if len(sys.argv) > 1 else 3
s = 4; N = 16
K = 2 * KH + 1
h, w = 40, 64          # H=160, W=256: tiles 5 x 4
H, W = h * s, w * s
rng = np.random.default_rng(1)
psf = o.gaussian_psf(K, 1.1)
shifts = np.array([[k % s, k // s] for k in range(N)], dtype=np.float64)  # one frame per phase
x = rng.random((1, H, W))
m = o.Model(s, psf, shifts)
truth = rng.random((H, W))
lr = np.stack([o.forward(m, k, truth) for k in range(N)])[:, None] + 0.01 * rng.standard_normal((N, 1, h, w))
obs_hr = o.upsample_observations(m, lr)
cost_ref, g_ref = o.data_term(m, x, obs_hr)

# yz from first principles: where does sample (k, q) land (transpose of the model without PSF)?
m0 = o.Model(s, None, shifts)
yz = np.zeros((H, W)); cnt = np.zeros((H, W))
for k in range(N):
    t = o.transpose(m0, k, lr[k, 0])
    ind = o.transpose(m0, k, np.ones((h, w)))
    yz += t; cnt += ind
# separable factors as factor_separable()
bi, bj = np.unravel_index(np.argmax(np.abs(psf)), psf.shape)
u = psf[:, bj].copy(); v = psf[bi, :] / psf[bi, bj]
assert np.allclose(np.outer(u, v), psf, atol=1e-17)

FT_W, TH, NT = 64, 32, 256
HB = KH; HX = max(KH + HB, 1); HXC = (HX + 1) & ~1
XOR_, XOC = HX - (KH + HB), HXC - (KH + HB)
XH, XW = TH + 2 * HX, FT_W + 2 * HXC
TW = FT_W + 2 * (KH + HB); TR = TH + 2 * HB; TP = TW | 1
BW = FT_W + 2 * HB; ZH = TH + 2 * KH; ZW = FT_W + 2 * KH; T2P = FT_W | 1
HYC = (KH + 1) & ~1; YW = FT_W + 2 * HYC; YH = TH + 2 * KH
A_D = max(XH * XW, TR * (BW | 1), ZH * T2P); B_D = max(TR * TP, ZH * (ZW | 1), 33 * 66)
assert YH * YW <= A_D and TR * TP <= B_D

def box(img, c0, r0, bw, bh):
    out = np.zeros((bh, bw))
    for r in range(bh):
        for c in range(bw):
            gr, gc = r0 + r, c0 + c
            if 0 <= gr < H and 0 <= gc < W: out[r, c] = img[gr, gc]
    return out.ravel().copy()

def slide(coef, L, nvalid, inp, outp):
    for l in range(L):
        if l < nvalid:
            outp(l, sum(coef[i] * inp(l + i) for i in range(K)))

def run_tile(tyi, txi):
    ty0, tx0 = tyi * TH, txi * FT_W
    A = np.full(A_D, np.nan); B = np.full(B_D, np.nan)
    A[:XH * XW] = box(x[0], tx0 - HXC, ty0 - HX, XW, XH)
    # 2a
    NSEG = max(NT // TW, 1); L = (TR + NSEG - 1) // NSEG
    for tid in range(NT):
        idv = tid
        while idv < TW * NSEG:
            c, seg = idv % TW, idv // TW; r0 = seg * L
            src = XOR_ * XW + XOC + r0 * XW + c; dst = r0 * TP + c
            slide(u, L, TR - r0, lambda i: A[src + i * XW], lambda l, val: B.__setitem__(dst + l * TP, val))
            idv += NT
    # y box -> A
    A[:] = np.nan
    A[:YH * YW] = box(yz, tx0 - HYC, ty0 - KH, YW, YH)
    # 2b in place, tails first (barrier), then all threads in arbitrary (here: reversed) order
    NSEG = max(NT // TR, 1); L = (BW + NSEG - 1) // NSEG
    assert TR * NSEG <= NT and L >= K - 1
    tails = {}
    for tid in range(NT):
        if tid < TR * NSEG:
            r, seg = tid % TR, tid // TR; c0 = seg * L; row = r * TP + c0
            tails[tid] = [B[row + L + i] if c0 + L + i < TW else 0.0 for i in range(K - 1)]
    for tid in reversed(range(NT)):
        if tid < TR * NSEG:
            r, seg = tid % TR, tid // TR; c0 = seg * L; row = r * TP + c0
            tl = tails[tid]
            # emulate sequential in-place semantic: reads of row[j] happen before the write of row[j]
            cache = {}
            def inp(j, row=row, tl=tl):
                return B[row + j] if j < L else tl[j - L]
            nvalid = BW - c0
            for l in range(L):
                if l < nvalid:
                    val = sum(v[i] * inp(l + i) for i in range(K))
                    B[row + l] = val
    # 3
    cost = 0.0
    ESEG = 4
    for tid in range(NT):
        c, rq = tid & 63, tid // 64
        for it in range((YH + ESEG - 1) // ESEG):
            r = rq + ESEG * it
            if r < YH:
                res = B[rq * TP + KH + c + it * ESEG * TP] - A[rq * YW + HYC + c + it * ESEG * YW]
                B[rq * TP + KH + c + it * ESEG * TP] = res
                if KH <= r < KH + TH: cost += res * res
        idv = tid
        while idv < 2 * KH * YH:
            r = idv // (2 * KH); hc = idv - r * 2 * KH
            cz = hc if hc < KH else hc + FT_W
            B[r * TP + cz] -= A[r * YW + cz - KH + HYC]
            idv += NT
    # 4a
    A[:] = np.nan
    NSEG = max(NT // ZH, 1); L = (FT_W + NSEG - 1) // NSEG
    for tid in range(NT):
        idv = tid
        while idv < ZH * NSEG:
            r, seg = idv % ZH, idv // ZH; c0 = seg * L
            src = r * TP + c0; dst = r * T2P + c0
            slide(u, L, FT_W - c0, lambda j: B[src + j], lambda l, val: A.__setitem__(dst + l, val))
            idv += NT
    # 4b
    g = np.zeros((TH, FT_W))
    EL = 8
    for tid in range(NT):
        ec, er0 = tid % FT_W, (tid // FT_W) * EL
        for l in range(EL):
            acc = sum(v[i] * A[(er0 + l + i) * T2P + ec] for i in range(K))
            g[er0 + l, ec] = 2.0 * s * s * acc
    return g, s * s * cost

for (tyi, txi) in [(2, 1), (1, 2), (3, 2)]:
    g, cost = run_tile(tyi, txi)
    ref = g_ref[0, tyi * TH:(tyi + 1) * TH, txi * FT_W:(txi + 1) * FT_W]
    assert not np.isnan(g).any()
    print("KH", KH, "tile", tyi, txi, "max abs diff", np.abs(g - ref).max(), "ref max", np.abs(ref).max(),
          "cnt==1 in Z region:", bool((cnt[tyi*TH-KH:(tyi+1)*TH+KH, txi*FT_W-KH:(txi+1)*FT_W+KH] == 1).all()))
