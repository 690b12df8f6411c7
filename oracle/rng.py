"""ORACLE (test infrastructure only) -- rand 0.8.5 `StdRng` (ChaCha12) and arkworks `UniformRand`.

Restates, from their published algorithms, the RNG crates the reference pins but does not vendor
(/root/reference/Cargo.lock:4315-4316 rand 0.8.5, :4346-4347 rand_chacha 0.3.1, :4375-4376
rand_core 0.6.4).  Reference call sites whose byte streams this reproduces:
  * core/src/sequencer/settlement/prover.rs:354   StdRng::seed_from_u64(inputs.batch_id) -> (r, s)
  * prover/src/bin/keygen.rs:87                   StdRng::seed_from_u64(0) -> trusted setup
  * prover/src/snarkjs.rs:153                     StdRng::seed_from_u64(42) -> setup + prove

Parity status: PINNED end-to-end by tests/test_oracle_kat.py::test_square_circuit_seed42_reproduces_reference_fixture
(setup + prove from seed 42 reproduce vk_snarkjs.json / proof_for_onchain.json byte for byte), when that test passes.
"""
import struct

from .bn254 import P, R, MONT_R, G1, G2, B_G1, B_G2, G2_COFACTOR, fq_sqrt, f2_sqrt, f2_add, f2_mul, f2_sqr, f2_neg

_MASK32 = 0xFFFFFFFF
_MASK64 = 0xFFFFFFFFFFFFFFFF


def _rotl(v, c):
    return ((v << c) & _MASK32) | (v >> (32 - c))


def _chacha_block(key_words, counter, rounds=12):
    # rand_chacha: 64-bit block counter in words 12,13; 64-bit stream id (0) in words 14,15
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & _MASK32, (counter >> 32) & _MASK32, 0, 0]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & _MASK32
        x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & _MASK32
        x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & _MASK32
        x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & _MASK32
        x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & _MASK32 for i in range(16)]


def seed_from_u64(state):
    """rand_core 0.6.4 SeedableRng::seed_from_u64: PCG32 expansion of a u64 into a 32-byte seed."""
    MUL = 6364136223846793005
    INC = 11634580027462260723
    out = b""
    for _ in range(8):
        state = (state * MUL + INC) & _MASK64
        xorshifted = (((state >> 18) ^ state) >> 27) & _MASK32
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & _MASK32 if rot else xorshifted
        out += struct.pack("<I", x)
    return out


class StdRng:
    """rand 0.8.5 StdRng = ChaCha12Rng over a BlockRng with a 64-word (4-block) buffer."""

    def __init__(self, seed32):
        assert len(seed32) == 32
        self.key = struct.unpack("<8I", seed32)
        self.counter = 0
        self.buf = []
        self.index = 64  # empty

    @classmethod
    def seed_from_u64(cls, v):
        return cls(seed_from_u64(v))

    def _generate(self):
        buf = []
        for i in range(4):
            buf += _chacha_block(self.key, self.counter + i)
        self.counter += 4
        self.buf = buf

    def next_u32(self):
        if self.index >= 64:
            self._generate()
            self.index = 0
        v = self.buf[self.index]
        self.index += 1
        return v

    def next_u64(self):
        # rand_core BlockRng::next_u64
        idx = self.index
        if idx < 63:
            self.index += 2
            return (self.buf[idx + 1] << 32) | self.buf[idx]
        if idx >= 64:
            self._generate()
            self.index = 2
            return (self.buf[1] << 32) | self.buf[0]
        lo = self.buf[63]
        self._generate()
        self.index = 1
        return (self.buf[0] << 32) | lo

    def gen_bool_standard(self):
        """rand 0.8 `Standard` for bool: sign bit of next_u32."""
        return bool(self.next_u32() >> 31)


def _rand_mont(rng, modulus):
    """ark-ff 0.5.0 `impl Distribution<Fp<MontBackend,4>> for Standard`: draw 4 u64 limbs (limb 0 first),
    clear the top 2 bits, interpret AS THE MONTGOMERY REPRESENTATION, reject if >= modulus.
    Returns the canonical value raw * R^{-1} mod modulus."""
    rinv = pow(MONT_R, -1, modulus)
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= _MASK64 >> 2
        raw = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        if raw < modulus:
            return raw * rinv % modulus


def rand_fr(rng):
    return _rand_mont(rng, R)


def rand_fq(rng):
    return _rand_mont(rng, P)


def rand_g1(rng):
    """ark-ec 0.5.0 `Distribution<Projective<P>> for Standard`: x <- Fq::rand, greatest <- bool,
    get_point_from_x_unchecked(x, greatest), then clear the cofactor (1 for G1)."""
    while True:
        x = rand_fq(rng)
        greatest = rng.gen_bool_standard()
        y = fq_sqrt((x * x * x + B_G1) % P)
        if y is None:
            continue
        ny = (-y) % P
        lo, hi = (y, ny) if y < ny else (ny, y)
        return (x, hi if greatest else lo)


def rand_g2(rng):
    while True:
        x = (rand_fq(rng), rand_fq(rng))
        greatest = rng.gen_bool_standard()
        y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B_G2))
        if y is None:
            continue
        ny = f2_neg(y)
        lo, hi = (y, ny) if (y[1], y[0]) < (ny[1], ny[0]) else (ny, y)
        pt = (x, hi if greatest else lo)
        return G2.mul(pt, G2_COFACTOR)
