"""Shared test fixtures: synthetic circuits, known-dlog bases, byte packing."""
import random

import numpy as np

from oracle import bn254 as bn
from oracle import groth16 as g16

R, P = bn.R, bn.P


def fr_bytes(vals):
    return b"".join(int(v % R).to_bytes(32, "little") for v in vals)


def fq_bytes(vals):
    return b"".join(int(v % P).to_bytes(32, "little") for v in vals)


def unpack32(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def g1_raw(points):
    return b"".join(bn.g1_to_raw(p) for p in points)


def g2_raw(points):
    return b"".join(bn.g2_to_raw(p) for p in points)


def arithmetic_bases(curve, gen, n, a, d):
    """P_i = (a + i d) G for i < n, via repeated addition + one batch inversion.  Returns (points, dlogs)."""
    F = curve
    cur = F.jac_mul(F.to_jac(gen), a % R)
    step = F.jac_mul(F.to_jac(gen), d % R)
    js, ks = [], []
    for i in range(n):
        js.append(cur)
        ks.append((a + i * d) % R)
        cur = F.jac_add(cur, step)
    return F.batch_to_affine(js), ks


def mimc7_chain(num_perm, seed, rounds=91):
    """Synthetic R1CS with the shape of forge/circuits/zelana_lib/src/poseidon.nr:30-46 (MiMC-7: 4 constraints
    per round: t = x + k + c_i ; t2 = t*t ; t4 = t2*t2 ; t6 = t4*t2 ; out = t6*t).  One public output.
    Returns (R1CS, z) with a satisfying assignment."""
    rnd = random.Random(seed)
    consts = [(i + 1) ** 3 + (i + 1) for i in range(rounds)]    # the reference's round constants (poseidon.nr:19-27)
    a, b, c = [], [], []
    # variables: 0 = ONE, 1 = public output ; witness from index 2
    z = [1, 0]
    x0 = rnd.randrange(R)
    key = rnd.randrange(R)
    z += [x0, key]
    xi, ki = 2, 3
    x = x0
    for p in range(num_perm):
        for i in range(rounds):
            t = (x + key + consts[i]) % R
            t2 = t * t % R
            t4 = t2 * t2 % R
            t6 = t4 * t2 % R
            t7 = t6 * t % R
            base = len(z)
            z += [t2, t4, t6, t7]
            lin_t = [(1, xi), (1, ki), (consts[i], 0)]
            a.append(list(lin_t)); b.append(list(lin_t)); c.append([(1, base)])
            a.append([(1, base)]); b.append([(1, base)]); c.append([(1, base + 1)])
            a.append([(1, base + 1)]); b.append([(1, base)]); c.append([(1, base + 2)])
            a.append([(1, base + 2)]); b.append(list(lin_t)); c.append([(1, base + 3)])
            x, xi = t7, base + 3
    # bind the public output: (x_final) * 1 = out
    a.append([(1, xi)]); b.append([(1, 0)]); c.append([(1, 1)])
    z[1] = x
    r1cs = g16.R1CS(num_instance=2, num_witness=len(z) - 2, a=a, b=b, c=c)
    return r1cs, z


def pk_parts(pk):
    """oracle ProvingKey -> kwargs for zelana_b200 Context.proving_key."""
    return dict(
        alpha_g1=bn.g1_to_raw(pk.vk.alpha_g1), beta_g1=bn.g1_to_raw(pk.beta_g1), beta_g2=bn.g2_to_raw(pk.vk.beta_g2),
        delta_g1=bn.g1_to_raw(pk.delta_g1), delta_g2=bn.g2_to_raw(pk.vk.delta_g2),
        a_query=g1_raw(pk.a_query), b_g1_query=g1_raw(pk.b_g1_query), b_g2_query=g2_raw(pk.b_g2_query),
        h_query=g1_raw(pk.h_query), l_query=g1_raw(pk.l_query))


def dot_mod_r(k, s):
    """sum_i k_i * s_i mod r for two [n, 8] u32 little-endian limb arrays, exactly: 16-bit limbs, float64 BLAS products over
    blocks of 2^18 rows (every partial sum < 2^50 is an exact double), recombined with Python integers."""
    n = k.shape[0]
    acc = [[0] * 16 for _ in range(16)]
    blk = 1 << 18
    for lo in range(0, n, blk):
        a = np.ascontiguousarray(k[lo:lo + blk]).view(np.uint16).astype(np.float64)   # [m, 16]
        b = np.ascontiguousarray(s[lo:lo + blk]).view(np.uint16).astype(np.float64)
        m = a.T @ b                                                                   # [16, 16], entries < 2^50
        for i in range(16):
            row = m[i]
            for j in range(16):
                acc[i][j] += int(row[j])
    total = 0
    for i in range(16):
        for j in range(16):
            total += acc[i][j] << (16 * (i + j))
    return total % R
