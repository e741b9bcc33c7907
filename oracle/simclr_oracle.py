"""oracle/simclr_oracle.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

float64 numpy restatement of the reference's dot-product InfoNCE ``SimCLRLoss`` (/root/reference/losses.py:162-202,
selected by ``main_mlp.py:146-147`` for ``--p 0``) and of the gradient autograd derives for it:

    neg_ik = z1_i . z3_k,  pos_i = z1_i . z2_i                           (losses.py:186-187; after the optional
                                                                          row normalisation of :179-184)
    lse_i  = logsumexp([neg_i: , pos_i] / tau)                            (:189-192)
    loss_i = 2 (alpha * (-pos_i / tau) + (1 - alpha) * lse_i)             (:191-197)
    returns mean loss, loss_i, [mean(-pos/tau), mean(lse)]                (:194-201)

Gradient of L = sum_i gl_i loss_i (gl = 1/B: the mean) with W = softmax over the B+1 logits of a row:
    dL/dz1_i = (2 gl_i / tau) [ -(alpha - (1-alpha) w+_i) z2_i + (1-alpha) sum_k W_ik z3_k ]
    dL/dz2_i = (2 gl_i / tau)   -(alpha - (1-alpha) w+_i) z1_i
    dL/dz3_k = (2 / tau) (1-alpha) sum_i gl_i W_ik z1_i
The normalisation (``normalize=True``) is NOT part of this function: callers normalise first (the product path does it
with torch ops as well and lets autograd chain through).

Pinned against outputs of the reference itself: tests/golden/simclr_*.npz (tests/golden/make_golden.py) via
tests/test_oracle_vs_golden.py."""
import numpy as np


def simclr(z1, z2, z3, tau=1.0, alpha=0.5, gl=None, need_grad=True, chunk=512):
    z1, z2, z3 = (np.asarray(a, dtype=np.float64) for a in (z1, z2, z3))
    B, d = z1.shape
    M = z3.shape[0]
    g = np.full(B, 1.0 / B) if gl is None else np.asarray(gl, dtype=np.float64)
    pos = (z1 * z2).sum(1)
    loss_i, lse = np.empty(B), np.empty(B)
    g1, g2, g3 = np.zeros((B, d)), np.zeros((B, d)), np.zeros((M, d))
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        logits = np.concatenate([z1[s:e] @ z3.T, pos[s:e, None]], axis=1) / tau
        m = logits.max(1, keepdims=True)
        ex = np.exp(logits - m)
        ssum = ex.sum(1, keepdims=True)
        lse[s:e] = (m + np.log(ssum))[:, 0]
        loss_i[s:e] = 2.0 * (alpha * (-pos[s:e] / tau) + (1.0 - alpha) * lse[s:e])
        if need_grad:
            W = ex / ssum
            wpos = W[:, M]
            cp = 2.0 * g[s:e] * (alpha - (1.0 - alpha) * wpos) / tau
            E = 2.0 * g[s:e] * (1.0 - alpha) / tau
            g1[s:e] = -cp[:, None] * z2[s:e] + E[:, None] * (W[:, :M] @ z3)
            g2[s:e] = -cp[:, None] * z1[s:e]
            g3 += (W[:, :M] * E[:, None]).T @ z1[s:e]
    out = dict(loss_mean=float(loss_i.mean()), loss_i=loss_i, lse=lse, pos=-pos,
               pos_mean=float((-pos / tau).mean()), neg_mean=float(lse.mean()))
    if need_grad:
        out.update(g1=g1, g2=g2, g3=g3)
    return out
