"""GPU diagnostic: sorted MSM entries against a numpy restatement of the signed-digit recoding; batched MSM against single MSMs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zelana_b200

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def rand_fr(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] %= 0x30644E72
    return a


def expected_entries(sc, c, nwin, table_n, first):
    """sc: [K, n, 8] u32 -> sorted list of (key, val)."""
    K, n, _ = sc.shape
    out = []
    nbuck = 1 << (c - 1)
    for p in range(K):
        for i in range(n):
            s = int.from_bytes(sc[p, i].tobytes(), "little")
            carry = 0
            for w in range(nwin):
                v = ((s >> (w * c)) & ((1 << c) - 1)) + carry
                neg = 0
                if v > nbuck:
                    v = (1 << c) - v
                    neg = 1
                    carry = 1
                else:
                    carry = 0
                if v:
                    out.append((p * nbuck + v - 1, (w * table_n + first + i) | (neg << 31)))
    return sorted(out)


def main():
    ctx = zelana_b200.Context(0)
    for n, K in ((300, 1), (300, 3), (3000, 9)):
        k = rand_fr(n, 81)
        bases = ctx.g1_bases_generate(torch.from_numpy(k.view(np.int32)).cuda(), n)
        c, nwin = bases.window()
        sc = rand_fr(n * K, 83).reshape(K, n, 8)
        sd = torch.from_numpy(sc.view(np.int32).copy()).cuda()
        keys, vals = ctx.debug_msm_entries(bases, sd, n, n, K)
        exp = expected_entries(sc, c, nwin, n, 0) if n <= 300 else None
        got = sorted(zip(keys.tolist(), vals.tolist()))
        print("n=%d K=%d c=%d nwin=%d entries=%d sorted_ok=%s" % (n, K, c, nwin, len(keys), bool(np.all(np.diff(keys.astype(np.int64)) >= 0))),
              "entries_match=%s" % (got == exp if exp is not None else "n/a"))
        single = [ctx.msm_g1(bases, sc[p]) for p in range(K)]
        gotb = ctx.debug_msm_batch(1, bases, sd, n, n, K)
        print("   batch == single per vector:", [a == b for a, b in zip(gotb, single)])
        if K > 1:
            same = np.repeat(sc[:1], K, axis=0).copy()
            gots = ctx.debug_msm_batch(1, bases, torch.from_numpy(same.view(np.int32).copy()).cuda(), n, n, K)
            print("   K identical vectors == single[0]:", [a == single[0] for a in gots])
        bases.free()
    ctx.close()


if __name__ == "__main__":
    main()
