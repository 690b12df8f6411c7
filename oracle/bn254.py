"""ORACLE (test infrastructure only) -- BN254 fields, groups, pairing, ark-serialize formats.

This file is a CPU restatement, in Python big-int arithmetic, of the arithmetic that the
reference prover delegates to arkworks 0.5.0 (ark-bn254 / ark-ff / ark-ec / ark-serialize;
pinned in /root/reference/Cargo.lock:226-229,290-293,344-347,497-500 and NOT vendored).
Call sites in the reference that this restates:
  * core/src/sequencer/settlement/prover.rs:263-277   ProvingKey/VerifyingKey::deserialize_compressed
  * core/src/sequencer/settlement/prover.rs:304-334   proof_to_solana_bytes (256 B LE, A negated)
  * prover/src/bin/convert_vk.rs:163-191              LE coordinate export, infinity = zeros
  * prover/src/snarkjs.rs:44-137                      compressed proof/VK export, snarkjs JSON (Fq2 as [c1,c0])
  * onchain-programs/verifier/.../lib.rs:497-547      Groth16 pairing check (used only as a KAT checker)

Parity status: PINNED for serialisation + pairing verification by the reference's committed
fixtures (tests/test_oracle_kat.py): prover/l2_vk.json, prover/l2_proof.json,
onchain-programs/verifier/proof_for_onchain.json + vk_snarkjs.json.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product path (zelana_b200/) never does.
"""

# ----------------------------------------------------------------------------- constants
# Fq modulus: onchain-programs/verifier/programs/onchain_verifier/src/lib.rs:9-10 (BE hex)
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
# Fr modulus: forge/crates/prover-worker/src/prover.rs:19-20 (decimal)
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
assert R == 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001

FR_GENERATOR = 5            # ark-bn254 Fr::GENERATOR (multiplicative generator, coset offset)
FR_TWO_ADICITY = 28
FR_ROOT_2_28 = pow(FR_GENERATOR, (R - 1) >> FR_TWO_ADICITY, R)
assert FR_ROOT_2_28 == 19103219067921713944291392827692070036145651957329286315305642004821462161904

MONT_R = 1 << 256           # arkworks MontBackend<_,4>: R = 2^256
B_G1 = 3
# G2 twist coefficient b' = 3/(9+u)
XI = (9, 1)


def inv_mod(a, m):
    return pow(a, -1, m)


# ----------------------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return ((a0 * b0 - a1 * b1) % P, (a0 * b1 + a1 * b0) % P)


def f2_sqr(a):
    a0, a1 = a
    return ((a0 + a1) * (a0 - a1) % P, 2 * a0 * a1 % P)


def f2_scalar(a, k):
    return (a[0] * k % P, a[1] * k % P)


def f2_inv(a):
    a0, a1 = a
    d = inv_mod((a0 * a0 + a1 * a1) % P, P)
    return (a0 * d % P, (-a1) * d % P)


def f2_conj(a):
    return (a[0], (-a[1]) % P)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


F2_ZERO = (0, 0)
F2_ONE = (1, 0)
B_G2 = f2_mul((3, 0), f2_inv(XI))
assert B_G2 == (
    19485874751759354771024239261021720505790618469301721065564631296452457478373,
    266929791119991161246907387137283842545076965332900288569378510910307636690,
)


def fq_sqrt(a):
    """p = 3 mod 4: candidate a^((p+1)/4); None if a is a non-residue."""
    a %= P
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a else None


def f2_sqrt(a):
    """Square root in Fq2 (complex method; any root -- callers choose the sign)."""
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        s = fq_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fq_sqrt((-a0) % P)
        return None if s is None else (0, s)
    norm = (a0 * a0 + a1 * a1) % P
    alpha = fq_sqrt(norm)
    if alpha is None:
        return None
    half = inv_mod(2, P)
    delta = (a0 + alpha) * half % P
    x0 = fq_sqrt(delta)
    if x0 is None:
        delta = (a0 - alpha) * half % P
        x0 = fq_sqrt(delta)
        if x0 is None:
            return None
    x1 = a1 * inv_mod(2 * x0 % P, P) % P
    r = (x0, x1)
    return r if f2_sqr(r) == (a0, a1) else None


# ----------------------------------------------------------------------------- generic short-Weierstrass
class Field:
    """Tiny vtable so one Jacobian implementation serves G1 (Fq ints) and G2 (Fq2 tuples)."""

    def __init__(self, add, sub, mul, sqr, neg, inv, zero, one, is_zero):
        self.add, self.sub, self.mul, self.sqr, self.neg, self.inv = add, sub, mul, sqr, neg, inv
        self.zero, self.one, self.is_zero = zero, one, is_zero


FQ = Field(
    lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P,
    lambda a: a * a % P, lambda a: (-a) % P, lambda a: inv_mod(a, P), 0, 1, lambda a: a % P == 0,
)
FQ2 = Field(f2_add, f2_sub, f2_mul, f2_sqr, f2_neg, f2_inv, F2_ZERO, F2_ONE, lambda a: a == F2_ZERO)

INF = None  # affine infinity


class Curve:
    """y^2 = x^3 + b over field F.  Affine points are (x, y) or None; Jacobian (X, Y, Z)."""

    def __init__(self, F, b, name):
        self.F, self.b, self.name = F, b, name

    def on_curve(self, pt):
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.sqr(y) == F.add(F.mul(F.sqr(x), x), self.b)

    def neg(self, pt):
        return None if pt is None else (pt[0], self.F.neg(pt[1]))

    # --- Jacobian arithmetic (a = 0)
    def to_jac(self, pt):
        F = self.F
        return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)

    def jac_zero(self):
        return (self.F.one, self.F.one, self.F.zero)

    def to_affine(self, j):
        F = self.F
        X, Y, Z = j
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jac_double(self, j):
        F = self.F
        X, Y, Z = j
        if F.is_zero(Z):
            return j
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(t, t)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C)
        C8 = F.add(C8, C8)
        C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        YZ = F.mul(Y, Z)
        return (X3, Y3, F.add(YZ, YZ))

    def jac_add(self, p, q):
        F = self.F
        X1, Y1, Z1 = p
        X2, Y2, Z2 = q
        if F.is_zero(Z1):
            return q
        if F.is_zero(Z2):
            return p
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jac_double(p)
            return self.jac_zero()
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def jac_add_affine(self, p, q):
        return p if q is None else self.jac_add(p, (q[0], q[1], self.F.one))

    def jac_neg(self, p):
        return (p[0], self.F.neg(p[1]), p[2])

    def jac_mul(self, p, k):
        if k < 0:
            return self.jac_mul(self.jac_neg(p), -k)
        acc = self.jac_zero()
        for bit in bin(k)[2:] if k else "":
            acc = self.jac_double(acc)
            if bit == "1":
                acc = self.jac_add(acc, p)
        return acc

    def mul(self, pt, k):
        return self.to_affine(self.jac_mul(self.to_jac(pt), k))

    def add(self, a, b):
        return self.to_affine(self.jac_add(self.to_jac(a), self.to_jac(b)))

    def msm_naive(self, bases, scalars):
        """Sum s_i * P_i by double-and-add; msm_bigint truncates to the shorter input
        (ark-ec 0.5.0 VariableBaseMSM::msm_bigint zips bases with scalars)."""
        acc = self.jac_zero()
        for pt, s in zip(bases, scalars):
            if s and pt is not None:
                acc = self.jac_add(acc, self.jac_mul(self.to_jac(pt), s))
        return self.to_affine(acc)

    def batch_to_affine(self, js):
        """Montgomery batch inversion of the Z coordinates."""
        F = self.F
        prods, acc = [], F.one
        for j in js:
            prods.append(acc)
            if not F.is_zero(j[2]):
                acc = F.mul(acc, j[2])
        inv = F.inv(acc)
        out = [None] * len(js)
        for i in range(len(js) - 1, -1, -1):
            X, Y, Z = js[i]
            if F.is_zero(Z):
                continue
            zi = F.mul(inv, prods[i])
            inv = F.mul(inv, Z)
            zi2 = F.sqr(zi)
            out[i] = (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))
        return out


G1 = Curve(FQ, B_G1, "G1")
G2 = Curve(FQ2, B_G2, "G2")
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
assert G1.on_curve(G1_GEN) and G2.on_curve(G2_GEN)
# G2 cofactor h2 = 2p - r (ark-bn254 g2::Config::COFACTOR)
G2_COFACTOR = 2 * P - R


# ----------------------------------------------------------------------------- Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)
# (u = w^6 - 9 satisfies u^2 = -1, and w^6 = 9 + u = XI: the sextic twist generator.)
def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                if bj:
                    t[i + j] += ai * bj
    for k in range(22, 11, -1):
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]


F12_ONE = [1] + [0] * 11


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


def _embed2(a, k):
    """Fq2 element a times w^k, k < 6, as an Fq12 coefficient list."""
    out = [0] * 12
    out[k] = (a[0] - 9 * a[1]) % P
    out[k + 6] = a[1] % P
    return out


ATE_LOOP = 29793968203157093288
_G12 = f2_pow(XI, (P - 1) // 3)   # xi^((p-1)/3)
_G13 = f2_pow(XI, (P - 1) // 2)   # xi^((p-1)/2)


def _g2_frobenius(q):
    return (f2_mul(f2_conj(q[0]), _G12), f2_mul(f2_conj(q[1]), _G13))


def _line(r, q, p):
    """Line through twisted points r, q (affine Fq2) evaluated at p in G1 -> (sparse Fq12, r+q)."""
    x1, y1 = r
    x2, y2 = q
    if x1 == x2 and y1 == y2:
        m = f2_mul(f2_scalar(f2_sqr(x1), 3), f2_inv(f2_scalar(y1, 2)))
    elif x1 == x2:
        raise ValueError("vertical line in Miller loop")
    else:
        m = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(m), x1), x2)
    y3 = f2_sub(f2_mul(m, f2_sub(x1, x3)), y1)
    # l = -yP + (m xP) w + (y1 - m x1) w^3
    l = [0] * 12
    l[0] = (-p[1]) % P
    a = _embed2(f2_scalar(m, p[0]), 1)
    b = _embed2(f2_sub(y1, f2_mul(m, x1)), 3)
    for i in range(12):
        l[i] = (l[i] + a[i] + b[i]) % P
    return l, (x3, y3)


def miller_loop(q, p):
    """Optimal-ate Miller loop for q in G2 (affine Fq2), p in G1 (affine). Infinity -> 1."""
    if q is None or p is None:
        return F12_ONE
    f = F12_ONE
    r = q
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        l, r2 = _line(r, r, p)
        f = f12_mul(f12_mul(f, f), l)
        r = r2
        if (ATE_LOOP >> i) & 1:
            l, r = _line(r, q, p)
            f = f12_mul(f, l)
    q1 = _g2_frobenius(q)
    nq2 = G2.neg(_g2_frobenius(q1))
    l, r = _line(r, q1, p)
    f = f12_mul(f, l)
    l, r = _line(r, nq2, p)
    f = f12_mul(f, l)
    return f


def final_exponentiation(f):
    return f12_pow(f, (P ** 12 - 1) // R)


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 for pairs [(g1_affine, g2_affine), ...]."""
    f = F12_ONE
    for p, q in pairs:
        f = f12_mul(f, miller_loop(q, p))
    return final_exponentiation(f) == F12_ONE


# ----------------------------------------------------------------------------- subgroup checks
def g1_in_subgroup(pt):
    return G1.on_curve(pt)  # cofactor 1


def g2_in_subgroup(pt):
    return G2.on_curve(pt) and G2.mul(pt, R) is None


# ----------------------------------------------------------------------------- ark-serialize 0.5.0 formats
# Layout confirmed on the reference fixtures: 32-byte little-endian canonical
# field elements; SW flags in the two top bits of the LAST byte: bit7 = "y is the larger of {y,-y}",
# bit6 = infinity.  Fq2 serialises c0 then c1; its ordering compares c1 first, then c0.
FLAG_NEG = 0x80
FLAG_INF = 0x40


def fq_to_bytes(x):
    return int(x % P).to_bytes(32, "little")


def fr_to_bytes(x):
    return int(x % R).to_bytes(32, "little")


def fr_from_bytes(b):
    return int.from_bytes(b, "little")


def _fq_is_larger(y):
    return y > (P - y) % P


def _f2_is_larger(y):
    ny = f2_neg(y)
    return (y[1], y[0]) > (ny[1], ny[0])


def g1_serialize(pt, compressed=True):
    if pt is None:
        n = 32 if compressed else 64
        out = bytearray(n)
        out[-1] |= FLAG_INF
        return bytes(out)
    x, y = pt
    flag = FLAG_NEG if _fq_is_larger(y) else 0
    if compressed:
        out = bytearray(fq_to_bytes(x))
    else:
        out = bytearray(fq_to_bytes(x) + fq_to_bytes(y))
    out[-1] |= flag
    return bytes(out)


def g2_serialize(pt, compressed=True):
    if pt is None:
        n = 64 if compressed else 128
        out = bytearray(n)
        out[-1] |= FLAG_INF
        return bytes(out)
    x, y = pt
    flag = FLAG_NEG if _f2_is_larger(y) else 0
    out = bytearray(fq_to_bytes(x[0]) + fq_to_bytes(x[1]))
    if not compressed:
        out += fq_to_bytes(y[0]) + fq_to_bytes(y[1])
    out[-1] |= flag
    return bytes(out)


class DecodeError(ValueError):
    pass


def _split_flags(b):
    b = bytearray(b)
    flags = b[-1] & 0xC0
    b[-1] &= 0x3F
    if flags == 0xC0:
        raise DecodeError("both SW flags set")
    return bytes(b), flags


def _fq_checked(b):
    v = int.from_bytes(b, "little")
    if v >= P:
        raise DecodeError("non-canonical Fq")
    return v


def g1_deserialize(b, compressed=True, validate=True):
    n = 32 if compressed else 64
    if len(b) != n:
        raise DecodeError("bad G1 length")
    body, flags = _split_flags(b)
    if flags & FLAG_INF:
        return None
    x = _fq_checked(body[:32])
    if compressed:
        y = fq_sqrt((x * x * x + B_G1) % P)
        if y is None:
            raise DecodeError("x not on G1")
        if _fq_is_larger(y) != bool(flags & FLAG_NEG):
            y = (-y) % P
    else:
        y = _fq_checked(body[32:64])
    pt = (x, y)
    if validate and not G1.on_curve(pt):
        raise DecodeError("G1 point not on curve")
    return pt


def g2_deserialize(b, compressed=True, validate=True):
    n = 64 if compressed else 128
    if len(b) != n:
        raise DecodeError("bad G2 length")
    body, flags = _split_flags(b)
    if flags & FLAG_INF:
        return None
    x = (_fq_checked(body[0:32]), _fq_checked(body[32:64]))
    if compressed:
        y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B_G2))
        if y is None:
            raise DecodeError("x not on G2")
        if _f2_is_larger(y) != bool(flags & FLAG_NEG):
            y = f2_neg(y)
    else:
        y = (_fq_checked(body[64:96]), _fq_checked(body[96:128]))
    pt = (x, y)
    if validate and not g2_in_subgroup(pt):
        raise DecodeError("G2 point not in the prime-order subgroup")
    return pt


# --- raw (flag-free) LE coordinate layouts used by the product's C ABI and by the reference's
#     "Solana" exporters (prover.rs:304-334, convert_vk.rs:163-191): infinity = all zero bytes.
def g1_to_raw(pt):
    return bytes(64) if pt is None else fq_to_bytes(pt[0]) + fq_to_bytes(pt[1])


def g2_to_raw(pt):
    if pt is None:
        return bytes(128)
    x, y = pt
    return fq_to_bytes(x[0]) + fq_to_bytes(x[1]) + fq_to_bytes(y[0]) + fq_to_bytes(y[1])


def g1_from_raw(b):
    if b == bytes(64):
        return None
    return (int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little"))


def g2_from_raw(b):
    if b == bytes(128):
        return None
    v = [int.from_bytes(b[i:i + 32], "little") for i in range(0, 128, 32)]
    return ((v[0], v[1]), (v[2], v[3]))
