"""Short-Weierstrass (a = 0) group arithmetic for BN254 / BLS12-381 G1 and G2.

Restates what the reference gets from ark-ec 0.2 (`GroupProjective` = Jacobian
X,Y,Z with identity Z = 0; `add_assign_mixed`, `double_in_place`, `into_affine`)
-- not in /root/reference (un-vendored crate).  Results are
compared as affine canonical coordinates, so the formula choice is immaterial.

Affine point: (x, y) or None (identity).  Jacobian: (X, Y, Z).
"""
from .fields import BLS12_381, BN254, FQ, FR, FpOps, Fp2Ops


class Curve:
    def __init__(self, name, F, b, gen, order, curve_id, group):
        self.name, self.F, self.b, self.gen, self.r = name, F, b, gen, order
        self.curve_id, self.group = curve_id, group

    # -- predicates ---------------------------------------------------------
    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sqr(y) == F.add(F.mul(F.sqr(x), x), self.b)

    # -- Jacobian -----------------------------------------------------------
    def identity(self):
        return (self.F.one, self.F.one, self.F.zero)

    def from_affine(self, P):
        if P is None:
            return self.identity()
        return (P[0], P[1], self.F.one)

    def to_affine(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def dbl(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z):
            return P
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        D = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C)
        C8 = F.add(C8, C8)
        C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def add(self, P, Q):
        F = self.F
        if F.is_zero(P[2]):
            return Q
        if F.is_zero(Q[2]):
            return P
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.dbl(P)
            return self.identity()
        H = F.sub(U2, U1)
        R = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(R), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def add_mixed(self, P, Qa):
        """P (Jacobian) + Qa (affine or None) -- ark `add_assign_mixed`."""
        if Qa is None:
            return P
        return self.add(P, (Qa[0], Qa[1], self.F.one))

    def neg(self, P):
        return (P[0], self.F.neg(P[1]), P[2])

    def neg_affine(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def mul(self, P, k):
        """k * P (Jacobian), plain double-and-add (k >= 0)."""
        R = self.identity()
        for i in reversed(range(k.bit_length())):
            R = self.dbl(R)
            if (k >> i) & 1:
                R = self.add(R, P)
        return R

    def mul_affine(self, Pa, k):
        return self.to_affine(self.mul(self.from_affine(Pa), k % self.r))

    def eq(self, P, Q):
        return self.to_affine(P) == self.to_affine(Q)

    def batch_to_affine(self, pts):
        """Montgomery-trick normalisation (ark `batch_normalization`)."""
        F = self.F
        prods, acc = [], F.one
        for P in pts:
            if not F.is_zero(P[2]):
                acc = F.mul(acc, P[2])
            prods.append(acc)
        inv = F.inv(acc)
        out = [None] * len(pts)
        for i in reversed(range(len(pts))):
            P = pts[i]
            if F.is_zero(P[2]):
                continue
            prev = prods[i - 1] if i > 0 else F.one
            zi = F.mul(inv, prev)
            inv = F.mul(inv, P[2])
            zi2 = F.sqr(zi)
            out[i] = (F.mul(P[0], zi2), F.mul(P[1], F.mul(zi2, zi)))
        return out

    # -- fixed-base helper (result-identical to ark FixedBaseMSM) -----------
    def fixed_base_table(self, Ga, bits, w=8):
        nwin = (bits + w - 1) // w
        table = []
        base = self.from_affine(Ga)
        for _ in range(nwin):
            row = [self.identity()]
            for d in range(1, 1 << w):
                row.append(self.add(row[-1], base))
            table.append(self.batch_to_affine(row))
            for _ in range(w):
                base = self.dbl(base)
        return (table, w)

    def fixed_base_mul(self, tbl, k):
        table, w = tbl
        R = self.identity()
        for j, row in enumerate(table):
            d = (k >> (w * j)) & ((1 << w) - 1)
            if d:
                R = self.add_mixed(R, row[d])
        return R


def _mk():
    curves = {}
    # ---- BLS12-381 ----
    p = FQ[BLS12_381].p
    r = FR[BLS12_381].p
    F1, F2 = FpOps(p), Fp2Ops(p)
    g1 = (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
          0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)
    g2 = ((0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
           0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
          (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
           0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE))
    curves[(BLS12_381, 1)] = Curve("bls12_381_g1", F1, 4, g1, r, BLS12_381, 1)
    curves[(BLS12_381, 2)] = Curve("bls12_381_g2", F2, (4, 4), g2, r, BLS12_381, 2)
    # ---- BN254 ----
    p = FQ[BN254].p
    r = FR[BN254].p
    F1, F2 = FpOps(p), Fp2Ops(p)
    b2 = F2.mul((3, 0), F2.inv((9, 1)))
    g2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))
    curves[(BN254, 1)] = Curve("bn254_g1", F1, 3, (1, 2), r, BN254, 1)
    curves[(BN254, 2)] = Curve("bn254_g2", F2, b2, g2, r, BN254, 2)
    return curves


CURVES = _mk()


def G1(curve_id):
    return CURVES[(curve_id, 1)]


def G2(curve_id):
    return CURVES[(curve_id, 2)]
