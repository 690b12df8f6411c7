"""Executes the literal PTX carry chains of csrc/fp.cuh on Python ints (tools/ptx_sim.py) against big-int math."""
import os
import random

import pytest

from conftest import ROOT
from tools import ptx_sim
from oracle.bn254 import P, R

SRC = open(os.path.join(ROOT, "zelana_b200", "csrc", "fp.cuh")).read()
BLOCKS = ptx_sim.extract_asm_blocks(SRC)
# file order: reduce_once, add, sub(2), modulus_minus, mont_step first, mont_step next, mont_step reduce, final merge
B_REDUCE, B_ADD, B_SUB1, B_SUB2, B_NEG, B_FIRST, B_NEXT, B_RED, B_MERGE, B_PROD_ADD = BLOCKS[:10]


def words(x):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def unwords(w):
    return sum(v << (32 * i) for i, v in enumerate(w))


def cfg(mod):
    env = {"C::M%d" % i: w for i, w in enumerate(words(mod))}
    env["C::INV"] = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    return env


def reduce_once(a, mod):
    env = cfg(mod)
    env.update({"a.v[%d]" % i: w for i, w in enumerate(words(a))})
    out = ptx_sim.run_asm(*B_REDUCE, env)
    t = unwords([out["t%d" % i] for i in range(8)])
    return t if out["brw"] == 0 else a


def mont_mul(a, b, mod):
    c = cfg(mod)
    aw, bw = words(a), words(b)
    E, O = [None] * 8, [None] * 8

    def step(first, X, Y, bi):
        env = dict(c)
        env.update({"a[%d]" % i: aw[i] for i in range(8)})
        env["b"] = bi
        if not first:
            env.update({"X[%d]" % i: X[i] for i in range(8)})
            env.update({"Y[%d]" % i: Y[i] for i in range(8)})
        out = ptx_sim.run_asm(*(B_FIRST if first else B_NEXT), env)
        for i in range(8):
            X[i], Y[i] = out["X[%d]" % i], out["Y[%d]" % i]
        env = dict(c)
        env["m"] = (X[0] * c["C::INV"]) & 0xFFFFFFFF
        env.update({"X[%d]" % i: X[i] for i in range(8)})
        env.update({"Y[%d]" % i: Y[i] for i in range(8)})
        out = ptx_sim.run_asm(*B_RED, env)
        for i in range(8):
            X[i], Y[i] = out["X[%d]" % i], out["Y[%d]" % i]
        assert X[0] == 0

    step(True, E, O, bw[0])
    for i in range(1, 8):
        if i & 1:
            step(False, O, E, bw[i])
        else:
            step(False, E, O, bw[i])
    env = {"E[%d]" % i: E[i] for i in range(8)}
    env.update({"O[%d]" % i: O[i] for i in range(8)})
    out = ptx_sim.run_asm(*B_MERGE, env)
    r = unwords([out["r.v[%d]" % i] for i in range(8)])
    assert r < 2 * mod
    return reduce_once(r, mod)


def mont_mul_add_mul(a, b, c, d, mod):
    """Fp::mul_add_mul: every step = product row of a * b_i, product row of c * d_i on top, ONE reduction row."""
    cf = cfg(mod)
    aw, bw, cw, dw = words(a), words(b), words(c), words(d)
    E, O = [None] * 8, [None] * 8

    def step(first, X, Y, i):
        env = dict(cf)
        env.update({"a[%d]" % k: aw[k] for k in range(8)})
        env["b"] = bw[i]
        if not first:
            env.update({"X[%d]" % k: X[k] for k in range(8)})
            env.update({"Y[%d]" % k: Y[k] for k in range(8)})
        out = ptx_sim.run_asm(*(B_FIRST if first else B_NEXT), env)
        for k in range(8):
            X[k], Y[k] = out["X[%d]" % k], out["Y[%d]" % k]
        before = unwords(X) + (unwords(Y) << 32)
        env = {"a[%d]" % k: cw[k] for k in range(8)}
        env["b"] = dw[i]
        env.update({"X[%d]" % k: X[k] for k in range(8)})
        env.update({"Y[%d]" % k: Y[k] for k in range(8)})
        out = ptx_sim.run_asm(*B_PROD_ADD, env)
        for k in range(8):
            X[k], Y[k] = out["X[%d]" % k], out["Y[%d]" % k]
        assert unwords(X) + (unwords(Y) << 32) == before + c * dw[i], "a carry was lost in the second product row"
        env = dict(cf)
        env["m"] = (X[0] * cf["C::INV"]) & 0xFFFFFFFF
        env.update({"X[%d]" % k: X[k] for k in range(8)})
        env.update({"Y[%d]" % k: Y[k] for k in range(8)})
        before = unwords(X) + (unwords(Y) << 32)
        out = ptx_sim.run_asm(*B_RED, env)
        for k in range(8):
            X[k], Y[k] = out["X[%d]" % k], out["Y[%d]" % k]
        assert X[0] == 0 and unwords(X) + (unwords(Y) << 32) == before + env["m"] * mod, "a carry was lost in the reduction row"

    step(True, E, O, 0)
    for i in range(1, 8):
        if i & 1:
            step(False, O, E, i)
        else:
            step(False, E, O, i)
    env = {"E[%d]" % i: E[i] for i in range(8)}
    env.update({"O[%d]" % i: O[i] for i in range(8)})
    out = ptx_sim.run_asm(*B_MERGE, env)
    r = unwords([out["r.v[%d]" % i] for i in range(8)])
    assert r == (unwords(O) >> 32) + unwords(E) < 3 * mod + 8, "merge overflowed 256 bits or the 3p bound fails"
    return reduce_once(reduce_once(r, mod), mod)


@pytest.mark.parametrize("mod", [P, R])
def test_mul_add_mul_ptx(mod):
    """The lazy-reduction product of the XYZZ formulas: (a b + c d) / R mod p with ONE Montgomery reduction, operands up to
    p itself (p - y for y = 0), every intermediate sum checked against exact integers."""
    rnd = random.Random(17)
    rinv = pow(1 << 256, -1, mod)
    edge = [0, 1, mod - 1, mod, (1 << 256) % mod, mod >> 1]
    cases = [(a, b, c, d) for a in edge for b in edge for c in (0, mod - 1, mod) for d in (1, mod - 1, mod)]
    cases += [tuple(rnd.randrange(mod) for _ in range(4)) for _ in range(300)]
    for a, b, c, d in cases:
        assert mont_mul_add_mul(a, b, c, d, mod) == (a * b + c * d) * rinv % mod, (a, b, c, d)


@pytest.mark.parametrize("mod", [P, R])
def test_mont_mul_ptx(mod):
    rnd = random.Random(7)
    rinv = pow(1 << 256, -1, mod)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, (1 << 512) % mod, (1 << 253), mod >> 1]
    cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(300)]
    for a, b in cases:
        assert mont_mul(a, b, mod) == a * b * rinv % mod


@pytest.mark.parametrize("mod", [P, R])
def test_add_sub_neg_ptx(mod):
    rnd = random.Random(9)
    c = cfg(mod)
    edge = [0, 1, mod - 1, mod >> 1, (mod >> 1) + 1]
    cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(300)]
    for a, b in cases:
        env = {"a.v[%d]" % i: w for i, w in enumerate(words(a))}
        env.update({"b.v[%d]" % i: w for i, w in enumerate(words(b))})
        out = ptx_sim.run_asm(*B_ADD, env)
        s = unwords([out["r.v[%d]" % i] for i in range(8)])
        assert reduce_once(s, mod) == (a + b) % mod
        out = ptx_sim.run_asm(*B_SUB1, env)
        d = [out["r.v[%d]" % i] for i in range(8)]
        brw = out["brw"]
        assert brw in (0, 0xFFFFFFFF)
        env2 = {"r.v[%d]" % i: d[i] for i in range(8)}
        env2.update({"C::M%d & brw" % i: w & brw for i, w in enumerate(words(mod))})
        out = ptx_sim.run_asm(*B_SUB2, env2)
        assert unwords([out["r.v[%d]" % i] for i in range(8)]) == (a - b) % mod
        if a:
            env3 = dict(c)
            env3.update({"a.v[%d]" % i: w for i, w in enumerate(words(a))})
            out = ptx_sim.run_asm(*B_NEG, env3)
            assert unwords([out["r.v[%d]" % i] for i in range(8)]) == mod - a
