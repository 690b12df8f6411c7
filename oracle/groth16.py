"""ORACLE (test infrastructure only) -- Groth16 over BN254 exactly as ark-groth16 0.5.0 computes it.

The reference enters this code at core/src/sequencer/settlement/prover.rs:408
(`Groth16::<Bn254>::prove(&self.proving_key, circuit, &mut rng)`), prover/src/bin/keygen.rs:87-91
(`circuit_specific_setup`) and prover/src/snarkjs.rs:156-163 (setup, prove, verify).  ark-groth16 /
ark-poly / ark-relations 0.5.0 are crates.io dependencies (Cargo.lock:411-414,440-443,473-476) that
are absent from /root/reference, so their published algorithms are restated here:
  * r1cs_to_qap.rs  LibsnarkReduction::{instance_map_with_evaluation, witness_map_from_matrices}
  * generator.rs    generate_parameters_with_qap
  * prover.rs       create_random_proof_with_reduction / create_proof_with_assignment / calculate_coeff
  * verifier.rs     prepare_inputs + 4-pairing product check
  * ark-poly        Radix2EvaluationDomain {fft, ifft, coset variants, evaluate_all_lagrange_coefficients}

Parity status: see oracle/rng.py header (pinned by the seed-42 SquareCircuit fixture when that test is green).
"""
from dataclasses import dataclass, field
from typing import List, Tuple

from .bn254 import (R, FR_GENERATOR, FR_TWO_ADICITY, FR_ROOT_2_28, G1, G2, G1_GEN, G2_GEN,
                    pairing_product_is_one, g1_serialize, g2_serialize, g1_deserialize, g2_deserialize,
                    g1_to_raw, g2_to_raw, fq_to_bytes, P)
from . import rng as orng


# ----------------------------------------------------------------------------- evaluation domain / NTT
def domain_size_for(n):
    """Radix2EvaluationDomain::new(n): next power of two (size 1 for n <= 1)."""
    size = 1
    while size < n:
        size <<= 1
    return size


def root_of_unity(size):
    log = size.bit_length() - 1
    assert 1 << log == size and log <= FR_TWO_ADICITY
    return pow(FR_ROOT_2_28, 1 << (FR_TWO_ADICITY - log), R)


def ntt(vals, omega):
    """In-order radix-2 DFT: out[k] = sum_j vals[j] omega^(jk).  Iterative, bit-reversal then DIT."""
    n = len(vals)
    if n == 1:
        return list(vals)
    log = n.bit_length() - 1
    a = [0] * n
    for i, v in enumerate(vals):
        a[int(format(i, "0%db" % log)[::-1], 2)] = v % R
    m = 1
    while m < n:
        w_m = pow(omega, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = w * a[k + j + m] % R
                u = a[k + j]
                a[k + j] = (u + t) % R
                a[k + j + m] = (u - t) % R
                w = w * w_m % R
        m <<= 1
    return a


def fft(vals, size=None):
    size = size or len(vals)
    v = list(vals) + [0] * (size - len(vals))
    return ntt(v, root_of_unity(size))


def ifft(vals):
    n = len(vals)
    ninv = pow(n, -1, R)
    out = ntt(vals, pow(root_of_unity(n), -1, R))
    return [x * ninv % R for x in out]


def coset_fft(vals, g=FR_GENERATOR):
    """domain.get_coset(g).fft_in_place: scale coefficient k by g^k, then FFT."""
    out, pw = [], 1
    for v in vals:
        out.append(v * pw % R)
        pw = pw * g % R
    return fft(out)


def coset_ifft(vals, g=FR_GENERATOR):
    """coset_domain.ifft_in_place: iFFT, then scale coefficient k by g^-k."""
    co = ifft(vals)
    ginv = pow(g, -1, R)
    out, pw = [], 1
    for v in co:
        out.append(v * pw % R)
        pw = pw * ginv % R
    return out


def evaluate_all_lagrange_coefficients(size, tau):
    """Radix2EvaluationDomain::evaluate_all_lagrange_coefficients (offset 1)."""
    omega = root_of_unity(size)
    z = (pow(tau, size, R) - 1) % R
    if z == 0:
        out, w = [], 1
        for _ in range(size):
            out.append(1 if w == tau % R else 0)
            w = w * omega % R
        return out
    # L_i(tau) = z * omega^i / (size * (tau - omega^i))
    sinv = pow(size, -1, R)
    out, w = [], 1
    for _ in range(size):
        out.append(z * w % R * sinv % R * pow((tau - w) % R, -1, R) % R)
        w = w * omega % R
    return out


# ----------------------------------------------------------------------------- R1CS
@dataclass
class R1CS:
    """ark-relations ConstraintMatrices after finalize(): rows of (coeff, variable) with variable
    index = instance index (0 is the constant ONE), or num_instance + witness index."""
    num_instance: int                      # includes the constant 1
    num_witness: int
    a: List[List[Tuple[int, int]]] = field(default_factory=list)
    b: List[List[Tuple[int, int]]] = field(default_factory=list)
    c: List[List[Tuple[int, int]]] = field(default_factory=list)

    @property
    def num_constraints(self):
        return len(self.a)

    def is_satisfied(self, z):
        ev = lambda row: sum(co * z[v] for co, v in row) % R
        return all(ev(ra) * ev(rb) % R == ev(rc) for ra, rb, rc in zip(self.a, self.b, self.c))


def _eval_row(row, z):
    return sum(co * z[v] for co, v in row) % R


def witness_map_from_matrices(r1cs, z):
    """LibsnarkReduction::witness_map_from_matrices -> h coefficients (domain_size of them)."""
    nc, ni = r1cs.num_constraints, r1cs.num_instance
    n = domain_size_for(nc + ni)
    a = [0] * n
    b = [0] * n
    for i in range(nc):
        a[i] = _eval_row(r1cs.a[i], z)
        b[i] = _eval_row(r1cs.b[i], z)
    for j in range(ni):
        a[nc + j] = z[j] % R
    a = coset_fft(ifft(a))
    b = coset_fft(ifft(b))
    ab = [x * y % R for x, y in zip(a, b)]
    c = [0] * n
    for i in range(nc):
        c[i] = _eval_row(r1cs.c[i], z)
    c = coset_fft(ifft(c))
    zinv = pow((pow(FR_GENERATOR, n, R) - 1) % R, -1, R)
    ab = [(x - y) * zinv % R for x, y in zip(ab, c)]
    return coset_ifft(ab)


# ----------------------------------------------------------------------------- keys
@dataclass
class VerifyingKey:
    alpha_g1: tuple
    beta_g2: tuple
    gamma_g2: tuple
    delta_g2: tuple
    gamma_abc_g1: list

    def serialize_compressed(self):
        out = g1_serialize(self.alpha_g1) + g2_serialize(self.beta_g2) + g2_serialize(self.gamma_g2)
        out += g2_serialize(self.delta_g2) + len(self.gamma_abc_g1).to_bytes(8, "little")
        return out + b"".join(g1_serialize(p) for p in self.gamma_abc_g1)

    @classmethod
    def deserialize_compressed(cls, b, validate=True):
        vk, off = cls._read(b, 0, validate)
        if off != len(b):
            raise ValueError("trailing bytes after VerifyingKey")
        return vk

    @classmethod
    def _read(cls, b, off, validate=True):
        al = g1_deserialize(b[off:off + 32], True, validate); off += 32
        g2s = []
        for _ in range(3):
            g2s.append(g2_deserialize(b[off:off + 64], True, validate)); off += 64
        n = int.from_bytes(b[off:off + 8], "little"); off += 8
        ic = []
        for _ in range(n):
            ic.append(g1_deserialize(b[off:off + 32], True, validate)); off += 32
        return cls(al, g2s[0], g2s[1], g2s[2], ic), off


@dataclass
class ProvingKey:
    vk: VerifyingKey
    beta_g1: tuple
    delta_g1: tuple
    a_query: list
    b_g1_query: list
    b_g2_query: list
    h_query: list
    l_query: list

    def serialize_compressed(self):
        """ark-groth16 ProvingKey field order: vk, beta_g1, delta_g1, a_query, b_g1_query, b_g2_query,
        h_query, l_query; Vec<T> = u64 LE length || elements."""
        def vec1(v):
            return len(v).to_bytes(8, "little") + b"".join(g1_serialize(p) for p in v)
        out = self.vk.serialize_compressed() + g1_serialize(self.beta_g1) + g1_serialize(self.delta_g1)
        out += vec1(self.a_query) + vec1(self.b_g1_query)
        out += len(self.b_g2_query).to_bytes(8, "little") + b"".join(g2_serialize(p) for p in self.b_g2_query)
        return out + vec1(self.h_query) + vec1(self.l_query)

    @classmethod
    def deserialize_compressed(cls, b, validate=True):
        vk, off = VerifyingKey._read(b, 0, validate)

        def g1(off):
            return g1_deserialize(b[off:off + 32], True, validate), off + 32

        def vec(off, size, fn):
            n = int.from_bytes(b[off:off + 8], "little"); off += 8
            out = []
            for _ in range(n):
                out.append(fn(b[off:off + size], True, validate)); off += size
            return out, off
        beta_g1, off = g1(off)
        delta_g1, off = g1(off)
        a_q, off = vec(off, 32, g1_deserialize)
        b1_q, off = vec(off, 32, g1_deserialize)
        b2_q, off = vec(off, 64, g2_deserialize)
        h_q, off = vec(off, 32, g1_deserialize)
        l_q, off = vec(off, 32, g1_deserialize)
        if off != len(b):
            raise ValueError("trailing bytes after ProvingKey")
        return cls(vk, beta_g1, delta_g1, a_q, b1_q, b2_q, h_q, l_q)


@dataclass
class Proof:
    a: tuple
    b: tuple
    c: tuple

    def serialize_compressed(self):
        return g1_serialize(self.a) + g2_serialize(self.b) + g1_serialize(self.c)

    def serialize_uncompressed(self):
        return g1_serialize(self.a, False) + g2_serialize(self.b, False) + g1_serialize(self.c, False)

    @classmethod
    def deserialize_compressed(cls, b):
        return cls(g1_deserialize(b[:32]), g2_deserialize(b[32:96]), g1_deserialize(b[96:128]))

    def to_solana_bytes(self):
        """core/src/sequencer/settlement/prover.rs:304-334: -A || B || C, raw 32 B LE coords, no flags."""
        return g1_to_raw(G1.neg(self.a)) + g2_to_raw(self.b) + g1_to_raw(self.c)


# ----------------------------------------------------------------------------- setup
def instance_map_with_evaluation(r1cs, tau):
    nc, ni = r1cs.num_constraints, r1cs.num_instance
    n = domain_size_for(nc + ni)
    zt = (pow(tau, n, R) - 1) % R
    u = evaluate_all_lagrange_coefficients(n, tau)
    nv = ni + r1cs.num_witness
    a = [0] * nv
    b = [0] * nv
    c = [0] * nv
    for j in range(ni):
        a[j] = u[nc + j]
    for i in range(nc):
        ui = u[i]
        for co, v in r1cs.a[i]:
            a[v] = (a[v] + ui * co) % R
        for co, v in r1cs.b[i]:
            b[v] = (b[v] + ui * co) % R
        for co, v in r1cs.c[i]:
            c[v] = (c[v] + ui * co) % R
    return a, b, c, zt, n


def _batch_mul(curve, gen, scalars):
    """BatchMulPreprocessing::batch_mul result = [s_i * gen] as affine (zero scalar -> infinity).
    Uses an 8-bit fixed window table; only the (unique) affine results matter for parity."""
    w = 8
    nwin = (254 + w - 1) // w
    gj = curve.to_jac(gen)
    table = []
    base = gj
    for _ in range(nwin):
        row = [curve.jac_zero()]
        for _k in range((1 << w) - 1):
            row.append(curve.jac_add(row[-1], base))
        table.append(row)
        base = curve.jac_add(row[-1], base)          # 2^w * base
    out = []
    for s in scalars:
        s %= R
        acc = curve.jac_zero()
        i = 0
        while s:
            d = s & ((1 << w) - 1)
            if d:
                acc = curve.jac_add(acc, table[i][d])
            s >>= w
            i += 1
        out.append(acc)
    return curve.batch_to_affine(out)


def generate_parameters(r1cs, alpha, beta, gamma, delta, g1_gen, g2_gen, tau):
    """ark-groth16 generate_parameters_with_qap with every random value made explicit."""
    ni = r1cs.num_instance
    a, b, c, zt, n = instance_map_with_evaluation(r1cs, tau)
    ginv = pow(gamma, -1, R)
    dinv = pow(delta, -1, R)
    gamma_abc = [(beta * a[i] + alpha * b[i] + c[i]) * ginv % R for i in range(ni)]
    l = [(beta * a[i] + alpha * b[i] + c[i]) * dinv % R for i in range(ni, len(a))]
    h_scalars, tp = [], 1
    for _ in range(n - 1):
        h_scalars.append(zt * dinv % R * tp % R)
        tp = tp * tau % R
    na, nb, nh, nl = len(a), len(b), len(h_scalars), len(l)
    g1_all = _batch_mul(G1, g1_gen, a + b + h_scalars + l + gamma_abc + [alpha, beta, delta])
    a_q = g1_all[:na]
    b1_q = g1_all[na:na + nb]
    h_q = g1_all[na + nb:na + nb + nh]
    l_q = g1_all[na + nb + nh:na + nb + nh + nl]
    gabc = g1_all[na + nb + nh + nl:na + nb + nh + nl + ni]
    alpha_g1, beta_g1, delta_g1 = g1_all[-3:]
    g2_all = _batch_mul(G2, g2_gen, b + [beta, gamma, delta])
    b2_q = g2_all[:nb]
    beta_g2, gamma_g2, delta_g2 = g2_all[-3:]
    vk = VerifyingKey(alpha_g1, beta_g2, gamma_g2, delta_g2, gabc)
    return ProvingKey(vk, beta_g1, delta_g1, a_q, b1_q, b2_q, h_q, l_q)


def circuit_specific_setup(r1cs, rng):
    """Groth16::circuit_specific_setup -> generate_random_parameters_with_reduction: RNG draw order
    alpha, beta, gamma, delta, G1::rand, G2::rand, then tau = sample_element_outside_domain."""
    alpha = orng.rand_fr(rng)
    beta = orng.rand_fr(rng)
    gamma = orng.rand_fr(rng)
    delta = orng.rand_fr(rng)
    g1_gen = orng.rand_g1(rng)
    g2_gen = orng.rand_g2(rng)
    n = domain_size_for(r1cs.num_constraints + r1cs.num_instance)
    while True:
        tau = orng.rand_fr(rng)
        if (pow(tau, n, R) - 1) % R != 0:
            break
    return generate_parameters(r1cs, alpha, beta, gamma, delta, g1_gen, g2_gen, tau)


# ----------------------------------------------------------------------------- prove / verify
def create_proof_with_assignment(pk, r, s, h, z, num_instance, msm_g1=None, msm_g2=None):
    """ark-groth16 prover.rs create_proof_with_assignment; z = full assignment [1, inputs.., aux..]."""
    msm_g1 = msm_g1 or G1.msm_naive
    msm_g2 = msm_g2 or G2.msm_naive
    J1, J2 = G1.to_jac, G2.to_jac
    aux = z[num_instance:]
    assignment = z[1:]
    h_acc = msm_g1(pk.h_query, h)
    l_acc = msm_g1(pk.l_query, aux)
    rs_delta = G1.jac_mul(J1(pk.delta_g1), r * s % R)

    def coeff(curve, initial, query, vk_param, msm):
        res = curve.jac_add_affine(initial, query[0])
        res = curve.jac_add_affine(res, msm(query[1:], assignment))
        return curve.jac_add_affine(res, vk_param)

    g_a = coeff(G1, G1.jac_mul(J1(pk.delta_g1), r), pk.a_query, pk.vk.alpha_g1, msm_g1)
    if r % R != 0:
        g1_b = coeff(G1, G1.jac_mul(J1(pk.delta_g1), s), pk.b_g1_query, pk.beta_g1, msm_g1)
    else:
        g1_b = G1.jac_zero()
    g2_b = coeff(G2, G2.jac_mul(J2(pk.vk.delta_g2), s), pk.b_g2_query, pk.vk.beta_g2, msm_g2)
    g_c = G1.jac_mul(g_a, s)
    g_c = G1.jac_add(g_c, G1.jac_mul(g1_b, r))
    g_c = G1.jac_add(g_c, G1.jac_neg(rs_delta))
    g_c = G1.jac_add_affine(g_c, l_acc)
    g_c = G1.jac_add_affine(g_c, h_acc)
    return Proof(G1.to_affine(g_a), G2.to_affine(g2_b), G1.to_affine(g_c))


def prove_with_rs(pk, r1cs, z, r, s, **kw):
    h = witness_map_from_matrices(r1cs, z)
    return create_proof_with_assignment(pk, r, s, h, z, r1cs.num_instance, **kw)


def prove(pk, r1cs, z, rng, **kw):
    """Groth16::prove: r then s drawn from the caller's RNG (prover.rs:354,408 seeds it with batch_id)."""
    r = orng.rand_fr(rng)
    s = orng.rand_fr(rng)
    return prove_with_rs(pk, r1cs, z, r, s, **kw)


def verify(vk, public_inputs, proof):
    """Groth16::verify: e(A,B) = e(alpha,beta) e(sum x_i IC_i, gamma) e(C,delta)."""
    if len(public_inputs) + 1 != len(vk.gamma_abc_g1):
        raise ValueError("wrong number of public inputs")
    acc = G1.to_jac(vk.gamma_abc_g1[0])
    for x, ic in zip(public_inputs, vk.gamma_abc_g1[1:]):
        acc = G1.jac_add(acc, G1.jac_mul(G1.to_jac(ic), x % R))
    vkx = G1.to_affine(acc)
    return pairing_product_is_one([
        (proof.a, proof.b), (G1.neg(vk.alpha_g1), vk.beta_g2),
        (G1.neg(vkx), vk.gamma_g2), (G1.neg(proof.c), vk.delta_g2)])


# ----------------------------------------------------------------------------- circuits used as fixtures
def square_circuit(x):
    """prover/src/snarkjs.rs:14-31 SquareCircuit: witness x, instance y, x*x = x_sq, (x_sq - y)*1 = 0.
    Variable order: instance [ONE, y]; witness [x, x_sq]."""
    y = x * x % R
    r1cs = R1CS(num_instance=2, num_witness=2,
                a=[[(1, 2)], [(1, 3), (R - 1, 1)]],
                b=[[(1, 2)], [(1, 0)]],
                c=[[(1, 3)], []])
    z = [1, y, x % R, y]
    return r1cs, z
