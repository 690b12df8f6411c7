"""N > 1 host logic on CPU: world_size-2 gloo run of zelana_b200.multi.ShardedMsm with an oracle-backed engine
(the GPU engine needs a B200; the sharding, all-gather and combine order are the same code on both)."""
import os
import random
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleEngine:
    """Partial = the rank's MSM as a canonical affine point padded to 128 B; combine = oracle point additions."""
    partial_bytes = 128

    def __init__(self, bases_raw):
        from oracle import cpu as orc
        self.orc = orc
        self.bases = orc.G1Bases.from_raw(bases_raw, threads=2)

    def msm_partial(self, scalars_local, n):
        raw = self.bases.msm(bytes(scalars_local.numpy()), threads=2)
        return torch.frombuffer(bytearray(raw + bytes(64)), dtype=torch.uint8)

    def combine(self, parts, k):
        from oracle import bn254 as bn
        acc = None
        for i in range(k):
            acc = bn.G1.add(acc, bn.g1_from_raw(bytes(parts[i * 128:i * 128 + 64].numpy())))
        return bn.g1_to_raw(acc)


def _worker(rank, world, port, n, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import R, fr_bytes, g1_raw, arithmetic_bases
        from oracle import bn254 as bn
        from zelana_b200.multi import ShardedMsm, shard_range
        rnd = random.Random(seed)
        pts, ks = arithmetic_bases(bn.G1, bn.G1_GEN, n, 11, 7)
        sc = [rnd.randrange(R) for _ in range(n)]
        lo, cnt = shard_range(n, world, rank)
        eng = OracleEngine(g1_raw(pts[lo:lo + cnt]))
        sm = ShardedMsm(eng)
        local = torch.frombuffer(bytearray(fr_bytes(sc[lo:lo + cnt])), dtype=torch.uint8)
        out = sm.run(local, cnt)
        exp = bn.g1_to_raw(bn.G1.mul(bn.G1_GEN, sum(k * s for k, s in zip(ks, sc)) % R))
        q.put((rank, out == exp))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    from zelana_b200.multi import shard_range
    for n in (0, 1, 7, 8, 9, 1 << 24):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1


@pytest.mark.parametrize("n", [5, 300])
def test_sharded_msm_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 1234 + n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
