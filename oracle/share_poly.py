"""CPU ORACLE (test infrastructure, NOT product code) -- share.RecoverCommit restated.

Follows /root/reference/share/poly.go:449-476 (RecoverCommit) and :418-445 (xyCommit) step by step:
  sort shares by index, keep the first t, x_i = I_i + 1, y_i = V_i;
  for each i:  num = prod_{j != i} x_j,  den = prod_{j != i} (x_j - x_i),  Acc += (num/den) * y_i.
Group arithmetic is delegated to a curve module (oracle.bn254 or oracle.bls12381); scalars are mod the
group order.  PARITY: the reference has no KAT for RecoverCommit (share/poly_test.go has property tests
only) -- "parity unpinned by fixtures", the result is fixed by mathematics (sum of lambda_i * Y_i).
"""
from __future__ import annotations


def lagrange_at_zero(indices, order: int):
    """lambda_i for x_i = I_i + 1 (poly.go:461-471)."""
    xs = [(i + 1) % order for i in indices]
    out = []
    for i, xi in enumerate(xs):
        num, den = 1, 1
        for j, xj in enumerate(xs):
            if j == i:
                continue
            num = num * xj % order
            den = den * ((xj - xi) % order) % order
        out.append(num * pow(den, order - 2, order) % order)
    return out


def recover_commit(curve, shares, t: int):
    """shares: list of (I, point); returns the affine point sum lambda_i * V_i over the first t by index."""
    if len(shares) < t:
        raise ValueError("share: not enough good public shares to reconstruct secret commitment")   # poly.go:452
    chosen = sorted(shares, key=lambda s: s[0])[:t]
    lam = lagrange_at_zero([s[0] for s in chosen], curve.ORDER)
    acc = None
    for l, (_, v) in zip(lam, chosen):
        acc = curve.g1_add(acc, curve.g1_mul(l, v))
    return acc


def lagrange_basis_matrix(indices, order: int):
    """L[j][k] = coefficient k of the j-th Lagrange basis polynomial over x_i = I_i + 1 (share/poly.go:513-545 lagrangeBasis:
    prod_{m != j} (x - x_m) / (x_j - x_m)), by plain polynomial products like the reference."""
    xs = [(i + 1) % order for i in indices]
    out = []
    for j, xj in enumerate(xs):
        poly, den = [1], 1
        for m, xm in enumerate(xs):
            if m == j:
                continue
            nxt = [0] * (len(poly) + 1)
            for k, c in enumerate(poly):                 # poly * (x - x_m)
                nxt[k + 1] = (nxt[k + 1] + c) % order
                nxt[k] = (nxt[k] - c * xm) % order
            poly = nxt
            den = den * ((xj - xm) % order) % order
        inv = pow(den, order - 2, order)
        out.append([c * inv % order for c in poly])
    return out


def recover_pubpoly(curve, shares, t: int, add=None, mul=None):
    """share.RecoverPubPoly (poly.go:480-508): commits[k] = sum_j L_j[k] * V_j over the first t shares by index"""
    add = add or curve.g1_add
    mul = mul or curve.g1_mul
    if len(shares) < t:
        raise ValueError("share: not enough good public shares to reconstruct secret commitment")
    chosen = sorted(shares, key=lambda s: s[0])[:t]
    L = lagrange_basis_matrix([s[0] for s in chosen], curve.ORDER)
    commits = []
    for k in range(t):
        acc = None
        for j, (_, v) in enumerate(chosen):
            acc = add(acc, mul(L[j][k], v))
        commits.append(acc)
    return commits
