"""TEST INFRASTRUCTURE ONLY — CPU restatement of the O(n3) *moment* form of ThreeBodyInteration that
``torch_m3gnet_b200/csrc/threebody_moment.cu`` implements (reference nn/interaction.py:187-223, :353-382).

The reference sums, for every first bond j of centre atom i, over all partner bonds k != j inside the three-body
cutoff:   red_j[l,n] = c_j * sum_k Y_l(u_j . u_k) * b_k[l,n]     (b_k = chi_ln(r_k) fc(r_k) sigma[dst k][l,n]).
For l <= 2 the angular factor is a polynomial in the unit bond vectors, so the pair sum factorises into per-atom
moments (S, V, M below) that cost O(n3) to build and O(1) per bond to evaluate; the same holds for every adjoint,
including the reference's Legendre-backward quirk (d/dx P_2 -> go*(2x + x*go), quadratic in the upstream
gradient), which needs the second-order moments Q / AQ.

Only ``tests/`` may import this module.  Plain torch; dtype follows the inputs (float64 for checking the algebra,
float32 for the accumulation-order statement).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

Y0 = math.sqrt(1.0 / (4.0 * math.pi))
Y1 = math.sqrt(3.0 / (4.0 * math.pi))
Y2 = math.sqrt(5.0 / (4.0 * math.pi))
SYM = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]  # symmetric 3x3 components (xx, yy, zz, xy, xz, yz)
NN = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]   # unordered (n, n') pairs of the quadratic quirk term


def _quad(M6: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """u^T M u for M given by its 6 symmetric components (last dim)."""
    return (M6[..., 0] * u[0] * u[0] + M6[..., 1] * u[1] * u[1] + M6[..., 2] * u[2] * u[2]
            + 2.0 * (M6[..., 3] * u[0] * u[1] + M6[..., 4] * u[0] * u[2] + M6[..., 5] * u[1] * u[2]))


def _matvec(M6: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """M u (.., 3) for symmetric M given by 6 components."""
    x = M6[..., 0] * u[0] + M6[..., 3] * u[1] + M6[..., 4] * u[2]
    y = M6[..., 3] * u[0] + M6[..., 1] * u[1] + M6[..., 5] * u[2]
    z = M6[..., 4] * u[0] + M6[..., 5] * u[1] + M6[..., 2] * u[2]
    return torch.stack([x, y, z], dim=-1)


def _outer6(u: torch.Tensor) -> torch.Tensor:
    return torch.stack([u[a] * u[b] for a, b in SYM])


def forward_atom(u: torch.Tensor, c: torch.Tensor, b: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Moments and reduced features of ONE centre atom.  u (n,3) unit vectors, c (n) cutoff values of the bonds as
    first bond, b (n,3,3) = b_k[l][n].  Sums run over members in ascending order (the kernel's stated order)."""
    n = u.shape[0]
    S0 = torch.zeros(3, dtype=u.dtype)
    V1 = torch.zeros(3, 3, dtype=u.dtype)       # [n][xyz]
    M2 = torch.zeros(3, 6, dtype=u.dtype)       # [n][sym]
    for k in range(n):
        S0 = S0 + b[k, 0]
        V1 = V1 + b[k, 1][:, None] * u[k][None, :]
        M2 = M2 + b[k, 2][:, None] * _outer6(u[k])[None, :]
    S2 = M2[:, 0] + M2[:, 1] + M2[:, 2]         # trace: |u| = 1
    acc = torch.zeros(n, 3, 3, dtype=u.dtype)
    for j in range(n):
        acc[j, 0] = Y0 * (S0 - b[j, 0])
        acc[j, 1] = Y1 * (V1 @ u[j] - b[j, 1])
        acc[j, 2] = Y2 * (1.5 * _quad(M2, u[j]) - 0.5 * S2 - b[j, 2])
    return dict(S0=S0, V1=V1, M2=M2, S2=S2, acc=acc, red=acc * c[:, None, None])


def backward_atom(u, r, c, b, q, fwd):
    """Adjoint of ``forward_atom`` for upstream q (n,3,3) = dL/d red.  Returns g_b (n,3,3), g_c (n) and the geometric
    gradient per bond as (g_v (n,3), g_r (n)) with dL/dv_j = g_v_j + g_r_j u_j  (g_r excludes the fc' term)."""
    n = u.shape[0]
    dt = u.dtype
    a = q * c[:, None, None]                    # dL/d acc
    g_c = (q * fwd["acc"]).sum(dim=(1, 2))
    # moments of a (roles swapped: j is the SECOND bond of (k, j))
    A0 = torch.zeros(3, dtype=dt)
    AV1 = torch.zeros(3, 3, dtype=dt)
    AM2 = torch.zeros(3, 6, dtype=dt)
    Q = torch.zeros(6, 6, dtype=dt)             # [nn'][sym]: sum_k b_k[2,n] b_k[2,n'] u_k u_k^T
    AQ = torch.zeros(6, 6, dtype=dt)            # same with a
    for k in range(n):
        o6 = _outer6(u[k])
        A0 = A0 + a[k, 0]
        AV1 = AV1 + a[k, 1][:, None] * u[k][None, :]
        AM2 = AM2 + a[k, 2][:, None] * o6[None, :]
        for p, (n1, n2) in enumerate(NN):
            Q[p] = Q[p] + (b[k, 2, n1] * b[k, 2, n2]) * o6
            AQ[p] = AQ[p] + (a[k, 2, n1] * a[k, 2, n2]) * o6
    A2s = AM2[:, 0] + AM2[:, 1] + AM2[:, 2]
    g_b = torch.zeros(n, 3, 3, dtype=dt)
    g_v = torch.zeros(n, 3, dtype=dt)
    g_r = torch.zeros(n, dtype=dt)
    V1, M2 = fwd["V1"], fwd["M2"]
    w = torch.tensor([1.0 if n1 == n2 else 2.0 for n1, n2 in NN], dtype=dt)
    for j in range(n):
        uj = u[j]
        g_b[j, 0] = Y0 * (A0 - a[j, 0])
        g_b[j, 1] = Y1 * (AV1 @ uj - a[j, 1])
        g_b[j, 2] = Y2 * (1.5 * _quad(AM2, uj) - 0.5 * A2s - a[j, 2])
        # d cos terms: G = sum_k gcos(j,k) u_k, X = sum_k gcos(j,k) cos(j,k); self terms (k = j, cos = 1) removed
        # role A (j first bond):   go1 = Y1 a_j[1].b_k[1], go2 = Y2 a_j[2].b_k[2]
        # role B (j second bond):  go1 = Y1 a_k[1].b_j[1], go2 = Y2 a_k[2].b_j[2]
        # gcos = go1 + go2*(2x + x*go2)   (reference quirk Q3)
        G = torch.zeros(3, dtype=dt)
        X = torch.zeros((), dtype=dt)
        for (pa, pb, Vm, Mm, Qm) in ((a[j], b[j], V1, M2, Q), (b[j], a[j], AV1, AM2, AQ)):
            # linear l = 1:  sum_k go1 u_k = Y1 sum_n pa[1,n] (Vm[n] - pb[1,n] u_j)
            G = G + Y1 * ((pa[1][:, None] * Vm).sum(0) - (pa[1] * pb[1]).sum() * uj)
            X = X + Y1 * ((pa[1] * (Vm @ uj)).sum() - (pa[1] * pb[1]).sum())
            # linear-in-go l = 2:  sum_k 2 x go2 u_k = 2 Y2 sum_n pa[2,n] (Mm[n] u_j - pb[2,n] u_j)
            Mu = _matvec(Mm, uj)                 # (3 n, 3)
            G = G + 2.0 * Y2 * ((pa[2][:, None] * Mu).sum(0) - (pa[2] * pb[2]).sum() * uj)
            X = X + 2.0 * Y2 * ((pa[2] * _quad(Mm, uj)).sum() - (pa[2] * pb[2]).sum())
            # quadratic quirk:  sum_k x go2^2 u_k = Y2^2 sum_{nn'} pa_n pa_n' (Qm[nn'] u_j - pb_n pb_n' u_j)
            pp = torch.stack([pa[2, n1] * pa[2, n2] for n1, n2 in NN]) * w
            bb = torch.stack([pb[2, n1] * pb[2, n2] for n1, n2 in NN])
            Qu = _matvec(Qm, uj)                 # (6, 3)
            G = G + Y2 * Y2 * ((pp[:, None] * Qu).sum(0) - (pp * bb).sum() * uj)
            X = X + Y2 * Y2 * ((pp * _quad(Qm, uj)).sum() - (pp * bb).sum())
        g_v[j] = G / r[j]
        g_r[j] = -X / r[j]
    return g_b, g_c, g_v, g_r
