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
