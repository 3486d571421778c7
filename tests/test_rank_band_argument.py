"""The rounding band of the fused rank counts (csrc/rank_counts_tc.cu), checked on the CPU by emulation.

The kernel classifies a pair by its Gram-form distance D~ (3xTF32 products, fp32 soft-min: score.cuh) and hands it to
the exact direct-form evaluation only when a threshold lies within  m = S (2^-18 + 2^-21 sqrt(D~ V)),
S = |e|^2 + max_k |p_k|^2, V = spread of the prototypes under the soft-min weights.  The counts equal the direct
route's whenever |D~ - direct| <= m.  Here both evaluations are emulated in numpy (tf32 splits with round-to-nearest,
fp32 accumulation, fp32 soft-min; the direct form op for op in float32) on the data shapes of the GPU tests, including
the two that shaped the band: far-apart prototypes (the exponents' error is amplified by sqrt(D V)) and a common
offset removed by centring (fp32 centring itself costs the most here).  The largest deviation must stay below half the
band -- the margin the GPU tests assert on the kernel (observed there: <= 0.25)."""
import numpy as np
import pytest

F32 = np.float32


def _tf32(x):
    """round-to-nearest (ties away) to a 10-bit mantissa: the bits of cvt.rna.tf32.f32"""
    u = x.astype(F32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(F32)


def _gram_3xtf32(E, P):
    """e.p with the operands split hi + lo (tf32 each), lo*hi + hi*lo + hi*hi, fp32 accumulation over d."""
    eh = _tf32(E); el = _tf32(E - eh)
    ph = _tf32(P); pl = _tf32(P - ph)
    acc = np.zeros((E.shape[0], P.shape[0]), F32)
    for j in range(E.shape[1]):                                   # fp32 accumulate, one dimension at a time
        acc = acc + (el[:, j, None] * ph[None, :, j] + eh[:, j, None] * pl[None, :, j] + eh[:, j, None] * ph[None, :, j]).astype(F32)
    return acc


def _softmin_gram(G, e2, P):
    """score.cuh softmin_from_gram in float32: s = softmax(2 g_k - |p_k|^2), dist = e2 - 2 s.g + s'PP's; also V."""
    p2 = (P * P).sum(1, dtype=F32)
    PP = (P @ P.T).astype(F32)
    a = (F32(2.0) * G - p2[None, :]).astype(F32)
    a = a - a.max(1, keepdims=True)
    w = np.exp(a.astype(F32)).astype(F32)
    s = (w / w.sum(1, keepdims=True, dtype=F32)).astype(F32)
    t1 = (s * G).sum(1, dtype=F32)
    m2 = np.einsum("nk,kl,nl->n", s, PP, s).astype(F32)
    D = (e2 - F32(2.0) * t1 + m2).astype(F32)
    V = ((s * p2[None, :]).sum(1, dtype=F32) - m2).astype(F32)
    return D, V


def _direct(E, P):
    """direct.cuh pcd_direct in float32, op for op (sequential fmaf chains emulated by float32 accumulation)."""
    N, d = E.shape
    K = P.shape[0]
    dk = np.zeros((N, K), F32)
    for k in range(K):
        acc = np.zeros(N, F32)
        for j in range(d):
            df = (E[:, j] - P[k, j]).astype(F32)
            acc = (df * df + acc).astype(F32)
        dk[:, k] = acc
    if K == 1:
        return dk[:, 0]
    mn = dk.min(1, keepdims=True)
    w = np.exp((mn - dk).astype(F32)).astype(F32)
    inv = (F32(1.0) / w.sum(1, dtype=F32)).astype(F32)
    dist = np.zeros(N, F32)
    for j in range(d):
        m = np.zeros(N, F32)
        for k in range(K):
            m = ((w[:, k] * inv) * P[k, j] + m).astype(F32)
        r = (E[:, j] - m).astype(F32)
        dist = (r * r + dist).astype(F32)
    return dist


@pytest.mark.parametrize("K,d,kind", [(3, 64, "normal"), (4, 20, "normal"), (1, 64, "normal"), (4, 128, "far"),
                                      (3, 16, "offset"), (8, 20, "normal"), (2, 32, "far")])
def test_gram_form_stays_inside_half_the_band(K, d, kind):
    rng = np.random.default_rng(17 * K + d)
    N, Q = 6000, 6
    E = rng.normal(size=(N, d)).astype(F32)
    Pq = (E[rng.integers(0, N, Q)][:, None, :] + 0.5 * rng.normal(size=(Q, K, d))).astype(F32)
    if kind == "far":
        Pq *= F32(4.0)
    if kind == "offset":
        E += F32(10.0); Pq += F32(10.0)
    mu = E.mean(0, dtype=np.float64).astype(F32)
    Ec = (E - mu).astype(F32)                                   # the catalog image and the query image are centred
    e2 = (Ec * Ec).sum(1, dtype=F32)
    worst = 0.0
    for q in range(Q):
        Pc = (Pq[q] - mu).astype(F32)
        D, V = _softmin_gram(_gram_3xtf32(Ec, Pc), e2, Pc)
        exact = _direct(E, Pq[q])                               # the direct form works on the raw rows
        S = e2 + (Pc * Pc).sum(1).max()
        band = S * (2.0 ** -18 + 2.0 ** -21 * np.sqrt(np.maximum(D, 0) * np.maximum(V, 0)))
        worst = max(worst, float((np.abs(D.astype(np.float64) - exact) / band).max()))
    assert worst < 0.5, worst
