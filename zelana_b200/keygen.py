"""Host-side mirror of the reference's key generation tool (prover/src/bin/keygen.rs:81-131).

`circuit_specific_setup` draws the randomness exactly as `Groth16::<Bn254>::circuit_specific_setup(circuit, &mut rng)`
does with `StdRng::seed_from_u64(0)` (keygen.rs:87-91) -- alpha, beta, gamma, delta, G1::rand, G2::rand, then tau outside
the evaluation domain -- and hands it to the GPU (`zkb_setup`: Lagrange coefficients at tau, QAP column evaluations and
the fixed-base batch multiplications).  The results are serialised with ark-serialize's compressed encoding
(`pk.serialize_compressed`, keygen.rs:100,117), so the bytes can be compared with the reference's key files.

Only integer formatting and the handful of field operations needed to sample two curve points happen here in Python.
"""
from .prover import FQ_MODULUS as P, FR_MODULUS as R, StdRng, fr_rand

_M64 = 0xFFFFFFFFFFFFFFFF
_RINV_FQ = pow(1 << 256, -1, P)
G2_COFACTOR = 2 * P - R


# ----------------------------------------------------------------------------- the little field arithmetic sampling needs
def _fq_sqrt(a):
    s = pow(a % P, (P + 1) // 4, P)          # p = 3 mod 4
    return s if s * s % P == a % P else None


def _f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _f2_inv(a):
    d = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


B_G1 = 3
B_G2 = _f2_mul((3, 0), _f2_inv((9, 1)))       # y^2 = x^3 + 3 / (9 + u)


def _f2_sqrt(a):
    """Complex method; any root (the caller orders the two)."""
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        s = _fq_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = _fq_sqrt(-a0)
        return None if s is None else (0, s)
    alpha = _fq_sqrt(a0 * a0 + a1 * a1)
    if alpha is None:
        return None
    half = pow(2, -1, P)
    x0 = _fq_sqrt((a0 + alpha) * half)
    if x0 is None:
        x0 = _fq_sqrt((a0 - alpha) * half)
        if x0 is None:
            return None
    x1 = a1 * pow(2 * x0, -1, P) % P
    r = (x0, x1)
    return r if _f2_mul(r, r) == (a0, a1) else None


# ----------------------------------------------------------------------------- ark-ff / ark-ec UniformRand
def fq_rand(rng: StdRng) -> int:
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        raw = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | ((limbs[3] & (_M64 >> 2)) << 192)
        if raw < P:
            return raw * _RINV_FQ % P


def _bool_rand(rng: StdRng) -> bool:
    return bool(rng.next_u32() >> 31)            # rand 0.8 Standard for bool


def g1_rand(rng: StdRng):
    """`G1Projective::rand`: x <- Fq, greatest <- bool, the point with that x and the larger / smaller y (cofactor 1)."""
    while True:
        x = fq_rand(rng)
        greatest = _bool_rand(rng)
        y = _fq_sqrt(x * x * x + B_G1)
        if y is None:
            continue
        lo, hi = sorted((y, (-y) % P))
        return (x, hi if greatest else lo)


def g2_rand_uncleared(rng: StdRng):
    """`G2Projective::rand` before the cofactor is cleared (Fq2 ordering: c1 first, then c0)."""
    while True:
        x = (fq_rand(rng), fq_rand(rng))
        greatest = _bool_rand(rng)
        y = _f2_sqrt(tuple((u + v) % P for u, v in zip(_f2_mul(_f2_mul(x, x), x), B_G2)))
        if y is None:
            continue
        ny = ((-y[0]) % P, (-y[1]) % P)
        lo, hi = sorted((y, ny), key=lambda t: (t[1], t[0]))
        return (x, hi if greatest else lo)


# ----------------------------------------------------------------------------- raw <-> ark-serialize compressed
def _le(x):
    return int(x).to_bytes(32, "little")


def g1_raw(pt):
    return bytes(64) if pt is None else _le(pt[0]) + _le(pt[1])


def g2_raw(pt):
    return bytes(128) if pt is None else _le(pt[0][0]) + _le(pt[0][1]) + _le(pt[1][0]) + _le(pt[1][1])


def g1_compress(raw: bytes) -> bytes:
    """x with bit 7 of the last byte = "y is the larger of {y, -y}", bit 6 = infinity."""
    if raw == bytes(64):
        return bytes(31) + b"\x40"
    y = int.from_bytes(raw[32:], "little")
    out = bytearray(raw[:32])
    if y > (P - y) % P:
        out[31] |= 0x80
    return bytes(out)


def g2_compress(raw: bytes) -> bytes:
    if raw == bytes(128):
        return bytes(63) + b"\x40"
    y0, y1 = int.from_bytes(raw[64:96], "little"), int.from_bytes(raw[96:], "little")
    out = bytearray(raw[:64])
    if (y1, y0) > ((-y1) % P, (-y0) % P):
        out[63] |= 0x80
    return bytes(out)


def _vec(chunks):
    return len(chunks).to_bytes(8, "little") + b"".join(chunks)


def serialize_keys(k: dict):
    """raw setup output -> (ProvingKey compressed bytes, VerifyingKey compressed bytes), ark-groth16 field order."""
    def split(b, sz):
        return [b[i:i + sz] for i in range(0, len(b), sz)]
    vk = (g1_compress(k["alpha_g1"]) + g2_compress(k["beta_g2"]) + g2_compress(k["gamma_g2"]) + g2_compress(k["delta_g2"]) +
          _vec([g1_compress(p) for p in split(k["gamma_abc_g1"], 64)]))
    pk = (vk + g1_compress(k["beta_g1"]) + g1_compress(k["delta_g1"]) +
          _vec([g1_compress(p) for p in split(k["a_query"], 64)]) +
          _vec([g1_compress(p) for p in split(k["b_g1_query"], 64)]) +
          _vec([g2_compress(p) for p in split(k["b_g2_query"], 128)]) +
          _vec([g1_compress(p) for p in split(k["h_query"], 64)]) +
          _vec([g1_compress(p) for p in split(k["l_query"], 64)]))
    return pk, vk


# ----------------------------------------------------------------------------- keygen
def circuit_specific_setup(ctx, num_instance, num_witness, a, b, c, rng: StdRng):
    """Groth16::circuit_specific_setup: returns (pk_bytes, vk_bytes, raw) with raw = the uncompressed affine outputs of zkb_setup.
    RNG draw order as in ark-groth16: alpha, beta, gamma, delta, G1::rand, G2::rand, then tau (rejected while in the domain)."""
    alpha, beta, gamma, delta = (fr_rand(rng) for _ in range(4))
    g1 = g1_rand(rng)
    g2u = g2_rand_uncleared(rng)
    g2 = ctx.scalar_mul(2, g2_raw(g2u), G2_COFACTOR.to_bytes(32, "little"))   # clear the cofactor on the GPU
    nc = (len(a[0]) - 1) if isinstance(a, tuple) else len(a)
    n = 1
    while n < nc + num_instance:
        n <<= 1
    while True:
        tau = fr_rand(rng)
        if pow(tau, n, R) != 1:
            break
    raw = ctx.setup(num_instance, num_witness, a, b, c, alpha=alpha, beta=beta, gamma=gamma, delta=delta, tau=tau,
                    g1_generator=g1_raw(g1), g2_generator=g2)
    pk, vk = serialize_keys(raw)
    return pk, vk, raw
