"""Range-sharded MSM across the GPUs of one box.

north_star: "MSMs shard naturally across the 8 GPUs of one box, by scalar range per GPU with the partial G1/G2 sums
combined over NVLink via NCCL".  One process per GPU (torch.distributed); rank r owns the contiguous range
shard_range(n, world, r) of (base, scalar) pairs: its bases (and their window tables) are loaded once, like a
proving key.  Per MSM every rank runs the whole Pippenger pipeline on its range and produces ONE projective
partial sum (128 B in G1, 256 B in G2); the only exchange is an all-gather of those world x 128 B (NCCL over
NVLink on GPUs, gloo in the CPU tests) followed by a `world`-term point addition -- the data path has no other
collective, and NCCL has no curve-addition reduce op, hence gather + local add rather than all-reduce.

The engine argument hides where the arithmetic runs: GpuMsmEngine drives libzkb200.so; the CPU tests inject an
oracle-backed engine to exercise this host logic under gloo with world_size 2.
"""
import torch
import torch.distributed as dist


def shard_range(n, world, rank):
    """Contiguous split of n items over `world` ranks, remainder to the lowest ranks -> (start, count)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


class GpuMsmEngine:
    """G1 or G2 partial MSM + combine on one GPU through the C ABI (zkb_msm_g*_dev / zkb_msm_g*_combine)."""

    def __init__(self, ctx, bases, group=1):
        self.ctx, self.bases, self.group = ctx, bases, group
        self.partial_bytes = 128 if group == 1 else 256
        self.affine_bytes = 64 if group == 1 else 128
        self.device = torch.device("cuda", ctx.device)
        self._part = torch.zeros(self.partial_bytes, dtype=torch.uint8, device=self.device)
        self._out = torch.zeros(self.affine_bytes, dtype=torch.uint8, device=self.device)

    def _shares_torch_stream(self):
        return self.ctx.stream_handle is not None and self.ctx.stream_handle == torch.cuda.current_stream(self.device).cuda_stream

    def msm_partial(self, scalars_dev, n):
        """scalars_dev: CUDA tensor of n x 32 B canonical scalars for THIS rank's range -> partial-sum tensor.
        The library queues its kernels on the Context's stream; the collective that follows runs on torch's current stream.
        Unless the two are the same stream (Context(device, stream=torch stream), as bench.py does) they must be ordered here:
        scalars written by torch before the MSM reads them, the partial sum complete before NCCL gathers it."""
        shared = self._shares_torch_stream()
        if not shared:
            torch.cuda.current_stream(self.device).synchronize()
        fn = self.ctx.msm_g1_dev if self.group == 1 else self.ctx.msm_g2_dev
        fn(self.bases, scalars_dev, n, out_partial_dev=self._part)
        if not shared:
            self.ctx.synchronize()
        return self._part

    def combine(self, parts, k):
        shared = self._shares_torch_stream()
        if not shared:
            torch.cuda.current_stream(self.device).synchronize()   # the all-gather that filled `parts`
        fn = self.ctx.msm_g1_combine if self.group == 1 else self.ctx.msm_g2_combine
        fn(parts, k, self._out)
        if not shared:
            self.ctx.synchronize()
        return self._out


class ShardedMsm:
    """sum_i s_i P_i over all ranks' ranges; every rank ends up with the canonical affine result."""

    def __init__(self, engine, process_group=None):
        self.engine = engine
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self._parts = None

    def run(self, scalars_local, n_local):
        part = self.engine.msm_partial(scalars_local, n_local)
        if self.world == 1:
            return self.engine.combine(part, 1)
        if self._parts is None or self._parts.device != part.device:
            self._parts = torch.empty(self.world * self.engine.partial_bytes, dtype=torch.uint8, device=part.device)
        chunks = list(self._parts.view(self.world, self.engine.partial_bytes).unbind(0))
        if part.is_cuda:
            dist.all_gather_into_tensor(self._parts, part, group=self.pg)
        else:
            dist.all_gather(chunks, part, group=self.pg)
        return self.engine.combine(self._parts, self.world)


class ShardedProver:
    """ONE Groth16 proof over the GPUs of a process group: rank r holds shard r of the proving key
    (Context.proving_key_shard / zkb_pk_load_shard), runs the witness map (replicated: no exchange) and the five MSMs over
    its range (zkb_prove_partial); the 768-byte partial records are all-gathered over NCCL and every rank finishes the
    identical proof (zkb_prove_combine).  The full assignment z is host data every rank already has (same node)."""

    def __init__(self, ctx, pk_shard, r1cs, process_group=None):
        from .api import PROVE_PARTIAL_BYTES
        self.ctx, self.pk, self.r1cs, self.pg = ctx, pk_shard, r1cs, process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.nbytes = PROVE_PARTIAL_BYTES
        dev = torch.device("cuda", ctx.device)
        self._part = torch.zeros(self.nbytes, dtype=torch.uint8, device=dev)
        self._parts = torch.zeros(self.world * self.nbytes, dtype=torch.uint8, device=dev)

    def prove(self, z_bytes, r_bytes, s_bytes):
        self.ctx.prove_partial(self.pk, self.r1cs, z_bytes, r_bytes, s_bytes, self._part)
        if self.world == 1:
            return self.ctx.prove_combine(self._part, 1, r_bytes, s_bytes)
        dist.all_gather_into_tensor(self._parts, self._part, group=self.pg)
        torch.cuda.current_stream().synchronize()
        return self.ctx.prove_combine(self._parts, self.world, r_bytes, s_bytes)
