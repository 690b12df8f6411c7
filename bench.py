#!/usr/bin/env python
"""bench.py -- BN254 G1 MSM (2^24 points) on N B200s, the configuration BASELINE.json's metric is quoted on.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 24] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one variable-base MSM  sum_i s_i * P_i  over n = 2^log_n synthetic (base, scalar) pairs:
  * bases  P_i = [k_i] G, generated on the GPU from seeded k_i (they are the proving key: resident in HBM);
  * scalars s_i uniform below r (top 32-bit limb drawn below r's top limb), seeded, identical for every N.
N > 1: the n pairs are split into N contiguous ranges (north_star: "by scalar range per GPU"), each rank runs the
whole Pippenger pipeline on its range, the N projective partial sums (128 B each) are all-gathered over NCCL and
added by one small kernel -> strong scaling of one 2^24 MSM.

Printed JSON (one line, rank 0): value = ms per MSM (device time, max over ranks, inputs resident in HBM);
e2e = the same through the host-buffer API (scalars copied H2D from pinned memory inside the timed region, the
64-byte affine result copied back); roofline = accumulate kernel vs the MEASURED INT32 multiply peak of this GPU;
cpu_baseline = oracle/cpu_oracle.cpp (arkworks' msm_bigint_wnaf restated) on the host cores, bounded sample.

--impl reference: the reference's CPU algorithm (the C++ restatement -- the Rust reference cannot be built in this
image) timed on the host cores for the same metric; rank 0 only.
"""
import argparse
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before torch initialises CUDA: see zelana_b200/__init__.py
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R_TOP_LIMB = 0x30644E72          # top 32-bit limb of the Fr modulus
BLOCK_LOG = 20                   # synthetic data is generated in 2^20-element blocks seeded by block index
SEED_BASES = 0xBA5E0000
SEED_SCALARS = 0x5EED0000
# W_G1(n) = n * 16 windows * 10 modmul * 136 mul32 (c = 16 signed-digit Pippenger, XYZZ mixed add)
MUL32_PER_POINT = 16 * 10 * 136
METRIC = "BN254 G1 MSM 2^%d ms"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-log-n", type=int, default=0, help="0 = choose for ~10-30 s of CPU work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-l2", action="store_true", help="msm_g1 workload: skip the L2-circuit proofs/s leg (size sweeps)")
    ap.add_argument("--workload", default="msm_g1", choices=["msm_g1", "ntt", "msm_sweep", "prove", "prove_sharded"],
                    help="msm_g1 = the contract line (default); ntt / msm_sweep = BASELINE.json configs 3 / 2 as extra sweeps")
    ap.add_argument("--concurrency", type=int, default=1, help="prove workload: contexts (host threads + streams) per GPU")
    ap.add_argument("--batch", type=int, default=0, help="prove workload: independent proofs per step over all GPUs (0 = one)")
    ap.add_argument("--cpu-baseline-prove", action="store_true", help="prove workload: time the CPU restatement even for large circuits")
    ap.add_argument("--logs", default="", help="comma-separated log2 sizes for the sweeps")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ synthetic data
def rand_fr_block(torch, seed, block, device):
    """2^20 canonical Fr elements as int32 [2^20, 8] little-endian limbs; depends only on (seed, block)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed + block)
    x = torch.randint(0, 1 << 32, (1 << BLOCK_LOG, 8), dtype=torch.int64, device=device, generator=g)
    x[:, 7] %= R_TOP_LIMB
    return x.to(torch.int32)


def rand_fr_range(torch, seed, start, count, device):
    blk = 1 << BLOCK_LOG
    out = torch.empty((count, 8), dtype=torch.int32, device=device)
    pos = 0
    while pos < count:
        b, o = divmod(start + pos, blk)
        take = min(blk - o, count - pos)
        out[pos:pos + take] = rand_fr_block(torch, seed, b, device)[o:o + take]
        pos += take
    return out


def numpy_scalars(np, seed, n):
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    s = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    s[:, 7] %= R_TOP_LIMB
    return s


def dot_mod_r_gpu(torch, k, s):
    """sum_i k_i * s_i mod r for two int32 [n, 8] little-endian limb tensors on the GPU, exactly: 16-bit limbs, float64 matmuls
    over blocks of 2^18 rows (every partial sum < 2^50 is an exact double), recombined with Python integers on the host."""
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    acc = [[0] * 16 for _ in range(16)]
    blk = 1 << 18

    def limbs16(x):
        lo = (x & 0xFFFF).to(torch.float64)
        hi = ((x >> 16) & 0xFFFF).to(torch.float64)
        return torch.stack((lo, hi), dim=2).reshape(x.shape[0], 16)      # limb 2j = low half of word j, 2j+1 = high half

    for lo in range(0, k.shape[0], blk):
        m = (limbs16(k[lo:lo + blk]).T @ limbs16(s[lo:lo + blk])).cpu().tolist()
        for i in range(16):
            for j in range(16):
                acc[i][j] += int(m[i][j])
    total = 0
    for i in range(16):
        for j in range(16):
            total += acc[i][j] << (16 * (i + j))
    return total % R


G1_GENERATOR_RAW = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_msm_rate(orc, np, log_n, threads, repeats=1, bases=None, scalars=None):
    """Seconds per MSM of 2^log_n points with the C++ restatement of arkworks' msm_bigint (bases pre-parsed)."""
    n = 1 << log_n
    own = bases is None
    if own:
        bases = orc.G1Bases.arithmetic(0x1234567, n, threads)
        scalars = numpy_scalars(np, SEED_SCALARS + log_n, n)
    best, out = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = bases.msm(scalars, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    if own:
        bases.free()
    return best, out


def choose_cpu_sample(orc, np, threads, budget_s, hi=22, lo=16):
    t18, _ = cpu_msm_rate(orc, np, 18, threads)
    lg = 18
    while lg < hi and t18 * (1 << (lg + 1 - 18)) * 0.9 <= budget_s:
        lg += 1
    while lg > lo and t18 * (1 << (lg - 18)) > budget_s * 1.5:
        lg -= 1
    return lg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu as orc
    try:
        threads = len(os.sched_getaffinity(0))      # torchrun exports OMP_NUM_THREADS=1: ask the OS, not OpenMP
    except AttributeError:
        threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    lg = args.cpu_sample_log_n or choose_cpu_sample(orc, np, threads, budget_s=150.0 / max(total, 1))
    lg = min(lg, args.log_n)
    n = 1 << lg
    bases = orc.G1Bases.arithmetic(0x1234567, n, threads)
    scalars = numpy_scalars(np, SEED_SCALARS + lg, n)
    for _ in range(args.warmup):
        bases.msm(scalars, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bases.msm(scalars, threads=threads)
    per_step = (time.perf_counter() - t0) / max(args.steps, 1)
    scale = float(1 << (args.log_n - lg))
    ms_full = per_step * 1e3 * scale
    sample = ("each step = one MSM of 2^%d points ((k0+i)G bases pre-parsed to Montgomery form, uniform scalars), "
              "time scaled x%d to 2^%d points; C++ restatement of ark-ec msm_bigint_wnaf (window parallelism only), "
              "not arkworks itself (no Rust toolchain / un-vendored crates)" % (lg, int(scale), args.log_n))
    line = {
        "impl": "reference", "metric": METRIC % args.log_n, "value": ms_full, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery)", "data": "synthetic",
        "config": {"workload": "bn254_g1_msm", "log_n": args.log_n, "parallelism": "cpu x%d threads" % threads},
        "cpu_baseline": {"value": ms_full, "unit": "ms", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ms_full, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import zelana_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE=%d but --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from zelana_b200.multi import GpuMsmEngine, ShardedMsm, shard_range
    n = 1 << args.log_n
    lo, shard = shard_range(n, world, rank)
    # a real (non-default) stream shared by torch and the library: CUDA events recorded on it see our kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = zelana_b200.Context(local, stream=stream.cuda_stream)

    # ---- synthetic workload (identical for every N): bases = proving-key points resident in HBM
    k = rand_fr_range(torch, SEED_BASES, lo, shard, dev)
    bases = ctx.g1_bases_generate(k, shard)
    ctx.synchronize()
    scal = rand_fr_range(torch, SEED_SCALARS, lo, shard, dev)
    # the bases are [k_i] G with known k_i: the MSM must equal [sum k_i s_i mod r] G.  The exact dot product of this rank's range
    # now, the comparison after the timed region (expected point through zkb_scalar_mul: an independent double-and-add path).
    dot_local = dot_mod_r_gpu(torch, k, scal)
    del k
    msm_c, msm_nwin = bases.window()
    engine = GpuMsmEngine(ctx, bases, group=1)
    sharded = ShardedMsm(engine)          # partial MSM per rank -> NCCL all-gather of world x 128 B -> point sum
    out_aff = engine._out

    def step_device():
        if world == 1:
            ctx.msm_g1_dev(bases, scal, shard, out_affine_dev=out_aff)
        else:
            sharded.run(scal, shard)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- INT32 multiply peak of this GPU, measured now (roofline denominator)
    peak_wide, _ = ctx.int32_peak(0)
    peak_pair, _ = ctx.int32_peak(1)
    peak = max(peak_wide, peak_pair)

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.profile(True)
    ctx.profile_reset()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = ctx.launch_count() - launches0
    phases = ctx.profile_read()
    ctx.profile(False)
    result_hex = bytes(out_aff.cpu().numpy()).hex()
    # ---- the timed result is checked, not just printed: sum over ranks of the exact dot products -> expected point
    R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, dot_local)
        dot_all = sum(parts) % R_MOD
    else:
        dot_all = dot_local
    expected_hex = bytes(ctx.scalar_mul(1, G1_GENERATOR_RAW, dot_all.to_bytes(32, "little"))).hex()
    if expected_hex != result_hex:
        raise SystemExit("PARITY FAILURE: MSM result %s != [sum k_i s_i] G = %s" % (result_hex[:32], expected_hex[:32]))

    # ---- end to end through the host-buffer API: H2D scalars (pinned) + MSM + D2H result, every step
    e2e = None
    if not args.no_e2e:
        host = torch.empty((shard, 8), dtype=torch.int32, pin_memory=True)
        host.copy_(scal)
        torch.cuda.synchronize()
        host_np = host.numpy().view(np.uint8).reshape(-1)
        out_host = torch.empty(64, dtype=torch.uint8, pin_memory=True)

        part_dev = engine._part
        parts_dev = torch.empty(world * engine.partial_bytes, dtype=torch.uint8, device=dev)

        def step_e2e():
            if world == 1:
                return ctx.msm_g1(bases, host_np)      # the C-ABI call a user makes: zkb_msm_g1(host scalars) -> 64 B
            # N > 1: the same sliced upload pipeline per rank through zkb_msm_g1_partial (host scalars in, projective partial
            # sum out), NCCL all-gather of the N x 128 B partials, zkb_msm_g1_combine, 64 B back to the host
            ctx.msm_g1_partial(bases, host_np, part_dev)
            dist.all_gather_into_tensor(parts_dev, part_dev)
            ctx.msm_g1_combine(parts_dev, world, out_aff)
            out_host.copy_(out_aff, non_blocking=True)
            stream.synchronize()
            return bytes(out_host.numpy())

        for _ in range(max(1, min(args.warmup, 2))):
            r_e2e = step_e2e()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            r_e2e = step_e2e()
        f1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1) / args.steps)
        assert bytes(r_e2e).hex() == result_hex, "host-buffer path and device path disagree"
        e2e = {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": shard * 32 * world, "d2h_bytes_per_step": 64 * world}

    # ---- roofline of the dominant kernel (bucket accumulation), device time from CUDA events around it
    acc_ms, acc_spans = phases.get("msm_g1_accumulate", (0.0, 0))
    roof = None
    if acc_spans:
        t_acc = acc_ms / acc_spans * 1e-3
        achieved = MUL32_PER_POINT * shard / t_acc
        roof = {"bound": "int32_mul", "kernel": "msm_accumulate_kernel<Fq>", "achieved": achieved / 1e12, "peak": peak / 1e12,
                "unit": "Tmul32/s", "frac": achieved / peak, "traffic": accumulate_traffic(args.log_n, world),
                "kernel_ms": t_acc * 1e3,
                "peak_source": "measured live: zkb_bench_int32_peak (mad.wide.u32 %.2f, lo/hi pairs %.2f Tmul32/s)" % (
                    peak_wide / 1e12, peak_pair / 1e12),
                "algorithmic_mul32_per_launch": MUL32_PER_POINT * shard,
                "whole_msm_frac": MUL32_PER_POINT * shard / (ms_dev * 1e-3) / peak,
                "hbm_gather_gbs": (shard * 16 * 64 + shard * 16 * 8) / t_acc / 1e9}
    phase_ms = {k2: v[0] / args.steps for k2, v in phases.items()}

    # ---- CPU baseline beside it (rank 0, N = 1 only): same bases/scalars, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu as orc
        threads = orc.max_threads()
        lg = min(args.cpu_sample_log_n or choose_cpu_sample(orc, np, threads, budget_s=20.0), args.log_n)
        m = 1 << lg
        cb = orc.G1Bases.from_raw(bases.read(0, m), threads)
        sc_np = scal[:m].cpu().numpy().view(np.uint8).reshape(-1)
        t_cpu, cpu_out = cpu_msm_rate(orc, np, lg, threads, bases=cb, scalars=sc_np)
        cb.free()
        gpu_out = ctx.msm_g1(bases, sc_np)     # same sub-range through the GPU path: must be byte-identical
        if bytes(gpu_out) != bytes(cpu_out):
            raise SystemExit("PARITY FAILURE: GPU MSM != CPU restatement on the first 2^%d points" % lg)
        scale = float(1 << (args.log_n - lg))
        cpu = {"value": t_cpu * 1e3 * scale, "unit": "ms", "cores": threads, "kind": "port",
               "sample": "first 2^%d of the 2^%d (base, scalar) pairs, %.2f s, scaled x%d; C++ restatement of ark-ec "
                         "msm_bigint_wnaf (parallel over windows only); GPU result on the same sample byte-identical"
                         % (lg, args.log_n, t_cpu, int(scale))}

    proofs = None
    extras = {}
    if not args.no_e2e:
        bases.free()
        bases = None
        del scal
        if rank == 0 and world == 1 and args.log_n == 24:
            extras = compact_config_lines(np, torch, zelana_b200, ctx, stream, dev, peak)
        if not args.no_l2:
            proofs = l2_sized_proofs_per_s(np, torch, dist, zelana_b200, local, world, rank, peak, cpu_baseline=not args.no_cpu_baseline)

    if rank == 0:
        line = {
            "metric": METRIC % args.log_n, "value": ms_dev, "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery)", "data": "synthetic",
            "config": {"workload": "bn254_g1_msm", "log_n": args.log_n, "points_per_gpu": shard,
                       "parallelism": "range-sharded x%d + NCCL all-gather of partial sums" % world if world > 1 else "single GPU",
                       "msm_window_bits": msm_c, "msm_windows": msm_nwin,
                       "l2": "inputs larger than L2: %d window tables x %d MB resident per GPU (%d MB of key-derived points, gathered "
                             "at random) + %d MB of scalars per GPU" % (msm_nwin, shard * 64 >> 20, msm_nwin * shard * 64 >> 20,
                                                                          shard * 32 >> 20)},
            "points_per_s": n / (ms_dev * 1e-3),
            "e2e": e2e, "gpu_launches": launches, "roofline": roof, "phase_ms_per_step": phase_ms,
            "cpu_baseline": cpu, "clocks": clocks, "result": result_hex,
            "result_verified": "equals [sum_i k_i s_i mod r] G (bases are [k_i] G; exact integer dot product, expected point via zkb_scalar_mul)",
            "groth16_proofs_per_s": proofs,
        }
        line.update(extras)
        emit(line)
    if bases is not None:
        bases.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def compact_config_lines(np, torch, zelana_b200, ctx, stream, dev, peak_int):
    """BASELINE.json configs 3 and 4 inside the default run, so that they are driver-run numbers: the four Fr transforms at
    2^24 (device-resident, CUDA events; both the contractual HBM fraction and the binding INT32 fraction) and one full Groth16
    prove of the forge-sized synthetic circuit (2^21 constraints; synthetic key of random points: timing-equivalent)."""
    out = {}
    hbm, hbm_src = measured_hbm_gbs()

    def timed(fn, steps=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    lg = 24
    n = 1 << lg
    a = rand_fr_range(torch, 0xA77E0000 + lg, 0, n, dev)
    b = torch.empty_like(a)
    row = {"log_n": lg, "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src, "algorithmic_bytes": 64 * n,
           "algorithmic_mul32": (n // 2) * lg * 136}
    for name, inv, coset in (("ntt", False, False), ("intt", True, False), ("coset_ntt", False, True), ("coset_intt", True, True)):
        ms = timed(lambda: ctx.ntt_dev(a, b, lg, inverse=inv, coset=coset))
        row[name + "_ms"] = ms
        row[name + "_hbm_frac"] = 64.0 * n / (ms * 1e-3) / (hbm * 1e9)
        row[name + "_int32_frac"] = (n / 2) * lg * 136.0 / (ms * 1e-3) / peak_int
    out["ntt_2p24"] = row
    del a, b
    torch.cuda.empty_cache()
    lg = 21
    num_perm = ((1 << lg) - 8) // (4 * 91)
    ni, nw, (A, B, Cm), z = mimc_r1cs_numpy(np, num_perm, seed=0xF0 + lg)
    m = ctx.r1cs(ni, nw, A, B, Cm)
    nv, n = ni + nw, 1 << lg
    k_len = max(nv + 2, n - 1) + 8
    k = rand_fr_range(torch, SEED_BASES, 0, k_len, dev)
    pk = ctx.proving_key_synthetic(nv, nw, n - 1, k, k_len)
    ctx.synchronize()
    del k
    z_np = torch.from_numpy(z).pin_memory().numpy().reshape(-1)
    r_b, s_b = (0x1234567).to_bytes(32, "little"), (0x7654321).to_bytes(32, "little")
    for _ in range(3):
        ctx.prove(pk, m, z_np, r_b, s_b)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        proof = ctx.prove(pk, m, z_np, r_b, s_b)
    ms = (time.perf_counter() - t0) / reps * 1e3
    out["prove_2p21"] = {"ms_per_proof": ms, "proofs_per_s": 1e3 / ms, "constraints": int(len(A[0]) - 1), "variables": int(nv),
                         "through": "zkb_prove, host z (64 MB H2D inside the timed region), wall clock, one context",
                         "key": "random curve points generated on the GPU ([k_i]G: timing-equivalent; the byte-for-byte check against "
                                "the CPU restatement at this size is tests/test_gpu_parity_scale.py)",
                         "proof_a": bytes(proof[0]).hex()[:32]}
    pk.free()
    m.free()
    torch.cuda.empty_cache()
    return out


def mimc_r1cs_numpy(np, num_perm, seed, rounds=91):
    """Synthetic R1CS of forge/circuits/zelana_lib/src/poseidon.nr:19-46 (MiMC-7 with the reference's round constants
    (i+1)^3 + (i+1): 4 constraints per round), built
    with numpy index arithmetic (tests/helpers.py::mimc7_chain is the small, loop-built twin checked against the oracle).
    -> (num_instance, num_witness, (A, B, C) CSR triples, z as uint8[nv, 32])."""
    import random
    R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rnd = random.Random(seed)
    consts = [(i + 1) ** 3 + (i + 1) for i in range(rounds)]    # the reference's round constants (poseidon.nr:19-27)
    x0, key = rnd.randrange(R), rnd.randrange(R)
    G = num_perm * rounds
    nv = 4 + 4 * G
    z = np.zeros((nv, 32), dtype=np.uint8)

    def put(i, v):
        z[i] = np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8)

    put(0, 1); put(2, x0); put(3, key)
    x = x0
    for g in range(G):
        t = (x + key + consts[g % rounds]) % R
        t2 = t * t % R
        t4 = t2 * t2 % R
        t6 = t4 * t2 % R
        x = t6 * t % R
        b = 4 + 4 * g
        put(b, t2); put(b + 1, t4); put(b + 2, t6); put(b + 3, x)
    put(1, x)
    g = np.arange(G, dtype=np.int64)
    base = 4 + 4 * g
    xi = np.where(g == 0, 2, base - 1)
    one = np.zeros(32, dtype=np.uint8); one[0] = 1
    cbytes = np.stack([np.frombuffer(c.to_bytes(32, "little"), dtype=np.uint8) for c in consts])[g % rounds]   # [G, 32]
    ki = np.full(G, 3, dtype=np.int64)
    zero = np.zeros(G, dtype=np.int64)

    def assemble(cols_per_row, coeff_per_row, last_col):
        """cols_per_row: list (4 rows of the group) of lists of column arrays; coeff: same shape, None = 1."""
        counts = [len(r) for r in cols_per_row]
        per_group = sum(counts)
        col = np.empty((G, per_group), dtype=np.uint32)
        co = np.zeros((G, per_group, 32), dtype=np.uint8)
        k = 0
        for r, cs in zip(cols_per_row, coeff_per_row):
            for c_arr, cf in zip(r, cs):
                col[:, k] = c_arr
                if cf is None:
                    co[:, k, 0] = 1
                else:
                    co[:, k] = cf
                k += 1
        rp_group = np.concatenate([[0], np.cumsum(counts)])[:4]
        row_ptr = (np.arange(G, dtype=np.uint64)[:, None] * per_group + rp_group[None, :].astype(np.uint64)).reshape(-1)
        row_ptr = np.concatenate([row_ptr, np.array([G * per_group, G * per_group + 1], dtype=np.uint64)])
        col = np.concatenate([col.reshape(-1), np.array([last_col], dtype=np.uint32)])
        co = np.concatenate([co.reshape(-1, 32), one[None, :]]).reshape(-1)
        return row_ptr, col, co

    lin, linc = [xi, ki, zero], [None, None, cbytes]
    A = assemble([lin, [base], [base + 1], [base + 2]], [linc, [None], [None], [None]], int(4 + 4 * (G - 1) + 3))
    B = assemble([lin, [base], [base], lin], [linc, [None], [None], linc], 0)
    Cm = assemble([[base], [base + 1], [base + 2], [base + 3]], [[None]] * 4, 1)
    return 2, nv - 2, (A, B, Cm), z


def l2_real_mul32_per_proof(circ, raw, c_of):
    """Arithmetic one proof of this circuit really needs on our schedule, in 32x32->64 multiply-accumulates (the unit of the MSM
    roofline): per MSM  entries x 10 products (XYZZ mixed addition) + 2^(c-1) buckets x 2 full additions x 14
    products, a G2 product = 3 G1 products (Karatsuba); 7 NTTs of (n/2) log2 n butterflies; 136 mul32 per product.
    Entries = (finite bases) x windows x (1 - 2^-c): zero digits and infinity bases produce none."""
    nv = circ.num_instance + circ.num_witness
    n = 1
    while n < circ.num_constraints + circ.num_instance:
        n <<= 1
    lg = n.bit_length() - 1

    def finite(rawq, size):
        import numpy as _np
        a = _np.frombuffer(rawq, dtype=_np.uint8).reshape(-1, size)
        return int((a.any(axis=1)).sum())

    msms = [("a_query", 64, 1), ("b_g1_query", 64, 1), ("l_query", 64, 1), ("h_query", 64, 1), ("b_g2_query", 128, 3)]
    products = 0.0
    for name, size, weight in msms:
        pts = finite(raw[name], size)
        c = c_of(len(raw[name]) // size)
        nwin = (255 + c - 1) // c
        products += weight * (pts * nwin * (1.0 - 2.0 ** -c) * 10 + (1 << (c - 1)) * 2 * 14)
    products += 7 * (n // 2) * lg
    return products * 136.0


def msm_window_for(n, point_bytes):
    """zelana_b200/csrc/msm.cuh msm_choose_window without the memory cap (small keys)."""
    best, best_cost = 0, 0.0
    for c in range(10, 24):
        nwin = (255 + c - 1) // c
        cost = nwin * n + 4.0 * (1 << (c - 1))
        if not best or cost < best_cost:
            best, best_cost = c, cost
    return best


def l2_sized_proofs_per_s(np, torch, dist, zelana_b200, local, world, rank, peak_mul32, lanes=16, per_gpu=2048, timed_steps=20,
                          cpu_baseline=True):
    """The other half of BASELINE.json's metric ("Groth16 proofs/s (L2 batch circuit)"): every GPU proves independent batches of
    the reference's own L2BlockCircuit (prover/src/l2_circuit.rs; the L2BlockCircuit::dummy() shape keygen.rs fixes: 6415
    constraints, domain 2^13) through ONE C call per step, zkb_l2_batch_prove: per proof `BatchProver::prove` end to end --
    witness assignment on the host (Poseidon folds, comparison bits) by a pool of `lanes` threads into pinned memory,
    StdRng(batch_id) -> (r, s), GPU prove in sub-batches of 256 proofs with batched kernels (zkb_prove_batch_begin), Solana
    byte layout.  The key is a real one: keygen.rs's flow (StdRng seed 0) with the setup on the GPU.  Every proof has its own
    batch id, hence its own roots, assignment and randomness; no communication between GPUs.  Aggregate proofs/s, wall clock
    around `timed_steps` steps, max over ranks."""
    from zelana_b200 import l2_circuit as l2
    ctx = zelana_b200.Context(local)
    t0 = time.perf_counter()
    circ, pk_bytes, vk_bytes, _raw = l2.keygen(ctx)
    keygen_s = time.perf_counter() - t0
    pk = ctx.proving_key_compressed(pk_bytes, validate=False)
    prover = l2.L2BatchProver(ctx, circ, pk, lanes=lanes)

    def batch(bid):
        ckt = l2.L2BlockCircuit(transactions=[l2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 1 + bid % 1000)],
                                initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): bid}, batch_id=bid)
        return ckt.with_inputs(l2.satisfying_inputs(ckt))

    circuits = [batch(rank * per_gpu + i + 1) for i in range(per_gpu)]
    pack = prover.marshal(circuits)
    for _ in range(3):     # buffers are sized by the first sub-batch on each of the two device slots
        proofs = prover.prove_marshalled(pack)
    # the batched path against the one-proof path (zkb_l2_prove: direct launches of the single-proof kernels): same bytes
    single = l2.L2Prover(ctx, circ, pk, vk_bytes)
    for i in (0, per_gpu // 2, per_gpu - 1):
        if single.prove_circuit(circuits[i]).proof_bytes != proofs[i]:
            raise SystemExit("PARITY FAILURE: batched L2 proof %d != the same proof made alone" % i)
    single.m.free()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(timed_steps):
        proofs = prover.prove_marshalled(pack)       # returns when every proof of the step is back in host memory
    dt = (time.perf_counter() - t0) / timed_steps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    # BASELINE.json config 5 literally: ONE batch of 64 independent proofs over all GPUs (64 / world per rank), wall clock from
    # the call to the last proof back in host memory, max over ranks
    per64 = max(1, 64 // world)
    pack64 = prover.marshal(circuits[:per64])
    prover.prove_marshalled(pack64)
    if world > 1:
        dist.barrier()
    t64 = []
    for _ in range(5):
        t0 = time.perf_counter()
        prover.prove_marshalled(pack64)
        t64.append(time.perf_counter() - t0)
    dt64 = sorted(t64)[len(t64) // 2]
    if world > 1:
        t = torch.tensor([dt64], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt64 = float(t.item())
    # one proof of this run checked the slow way (rank 0): the assignment satisfies the matrices
    ok = None
    if rank == 0:
        ok = circ.is_satisfied(circ.assign(circuits[0]))[0]
        if not ok:
            raise SystemExit("PARITY FAILURE: L2 circuit assignment does not satisfy its own constraints")
    t1 = time.perf_counter()
    for _ in range(8):
        circ.assign(circuits[0])
    assign_ms = (time.perf_counter() - t1) / 8 * 1e3
    total = world * per_gpu
    # the CPU restatement of arkworks' prover beside it (rank 0, N = 1): same key, same assignment, same (r, s) -> same bytes
    cpu = None
    if cpu_baseline and rank == 0 and world == 1:
        from oracle import cpu as orc
        threads = orc.max_threads()
        cpk = orc.ProvingKey(**{k2: _raw[k2] for k2 in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "a_query",
                                                       "b_g1_query", "b_g2_query", "h_query", "l_query")})
        a, b, c = circ.matrices()
        cm = orc.R1cs(circ.num_instance, circ.num_witness, csr=(a, b, c))
        zb = circ.assign(circuits[0])
        r_b, s_b = l2.prover_randomness(circuits[0].batch_id)
        reps = 5
        t2 = time.perf_counter()
        for _ in range(reps):
            cproof = orc.prove(cpk, cm, zb, r_b, s_b, threads=threads)
        t_cpu = (time.perf_counter() - t2) / reps
        if zelana_b200.proof_to_solana_bytes(*cproof) != proofs[0]:
            raise SystemExit("PARITY FAILURE: L2 circuit proof from the GPU != CPU restatement's proof (same key, witness, r, s)")
        cpu = {"value": 1.0 / t_cpu, "unit": "proofs/s", "cores": threads, "kind": "port",
               "sample": "%d proofs of the same circuit, key and assignment, %.1f ms each; C++ restatement of ark-groth16 "
                         "(witness map + 5 MSMs), proof bytes identical to the GPU's" % (reps, t_cpu * 1e3)}
    # roofline of this leg: the multiply-accumulates one proof really needs on our schedule / the time it takes / the INT32 peak
    mul32 = l2_real_mul32_per_proof(circ, _raw, lambda cnt: msm_window_for(cnt + 2, 64))
    rate = per_gpu / dt                       # per GPU
    roof = {"bound": "int32_mul", "achieved": mul32 * rate / 1e12, "peak": peak_mul32 / 1e12, "unit": "Tmul32/s",
            "frac": mul32 * rate / peak_mul32, "traffic": None, "real_mul32_per_proof": mul32,
            "note": "REAL work (not the c = 16 normalisation of the MSM line): 4 G1 + 1 G2 batched MSMs at the key's window width "
                    "(entries x 10 products + buckets x 28), 7 NTTs of 2^13, 136 mul32 per product; wall clock includes the host "
                    "witness assignment, H2D of every assignment and D2H of every proof"}
    prover.close()
    pk.free()
    ctx.close()
    return {"value": total / dt, "unit": "proofs/s", "proofs": total, "proofs_per_gpu_per_step": per_gpu, "steps": timed_steps,
            "host_threads_per_gpu": lanes, "ms_per_batch": dt * 1e3,
            "host_assign_ms_per_proof": assign_ms, "host_cores": os.cpu_count(), "keygen_s": keygen_s, "assignment_satisfied": ok,
            "batched_equals_single": True, "roofline": roof, "cpu_baseline": cpu,
            "batch_of_64": {"ms": dt64 * 1e3, "proofs_per_s": per64 * world / dt64, "proofs_per_gpu": per64,
                            "what": "BASELINE.json config 5: one batch of 64 independent L2 proofs over the GPUs, median of 5"},
            "circuit": "L2BlockCircuit::dummy() shape (prover/src/l2_circuit.rs): %d constraints, %d witness variables, domain 2^13; "
                       "one zkb_l2_batch_prove call per step = %d x BatchProver::prove end to end (host witness assignment + GPU "
                       "prove + Solana bytes), %d host threads per GPU, sub-batches of 256 proofs through batched kernels, real key "
                       "from the GPU trusted setup (StdRng seed 0 as keygen.rs), independent proofs sharded over the GPUs with no "
                       "communication" % (circ.num_constraints, circ.num_witness, per_gpu, lanes)}


def run_prove(args):
    """BASELINE.json configs 4 and 5: full Groth16 proves (7 NTTs + 4 G1 MSMs + 1 G2 MSM each) of a synthetic MiMC circuit:
    --log-n 21 = forge-sized (config 4, default); --log-n 13 --batch 64 = "batch of independent L2-sized proofs", one proof
    per GPU at a time per context, proofs sharded over the ranks with no communication (config 5)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import zelana_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    K = max(1, args.concurrency)
    streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
    torch.cuda.set_stream(streams[0])
    ctxs = [zelana_b200.Context(local, stream=st.cuda_stream) for st in streams]
    ctx = ctxs[0]
    lg = args.log_n if args.log_n != 24 else 21
    num_perm = ((1 << lg) - 8) // (4 * 91)
    t0 = time.perf_counter()
    ni, nw, (A, B, Cm), z = mimc_r1cs_numpy(np, num_perm, seed=0xF0 + lg)
    t_build = time.perf_counter() - t0
    m = ctx.r1cs(ni, nw, A, B, Cm)
    assert m.log_domain == lg, (m.log_domain, lg)
    nv, n = ni + nw, 1 << lg
    # Small circuits: the key (random curve points generated on the GPU, read back) is loaded into BOTH provers, so the CPU
    # restatement's proof can be compared byte for byte; forge-sized: a device-only synthetic key.
    with_cpu = rank == 0 and world == 1 and K == 1 and not args.no_cpu_baseline and (lg <= 16 or args.cpu_baseline_prove)
    parts = None
    t_setup = 0.0
    t0 = time.perf_counter()
    if with_cpu:
        # a REAL key: Groth16 trusted setup on the GPU (zkb_setup) with StdRng(0), as prover/src/bin/keygen.rs:87 does
        from zelana_b200 import keygen as kg
        from zelana_b200.prover import StdRng
        _, _, raw = kg.circuit_specific_setup(ctx, ni, nw, A, B, Cm, StdRng.seed_from_u64(0)) if lg <= 16 else (None, None, None)
        if raw is None:   # large: skip the (Python) compressed serialisation, keep the raw affine outputs of zkb_setup
            rng0 = StdRng.seed_from_u64(0)
            from zelana_b200.prover import fr_rand
            al, be, ga, de = (fr_rand(rng0) for _ in range(4))
            g1 = kg.g1_rand(rng0)
            g2 = ctx.scalar_mul(2, kg.g2_raw(kg.g2_rand_uncleared(rng0)), kg.G2_COFACTOR.to_bytes(32, "little"))
            tau = fr_rand(rng0)
            raw = ctx.setup(ni, nw, A, B, Cm, alpha=al, beta=be, gamma=ga, delta=de, tau=tau, g1_generator=kg.g1_raw(g1),
                            g2_generator=g2)
        parts = {k2: raw[k2] for k2 in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "a_query", "b_g1_query",
                                        "b_g2_query", "h_query", "l_query")}
        t_setup = time.perf_counter() - t0
        pk = ctx.proving_key(**parts)
    else:
        k_len = max(nv + 2, n - 1) + 8
        k = rand_fr_range(torch, SEED_BASES, 0, k_len, dev)
        pk = ctx.proving_key_synthetic(nv, nw, n - 1, k, k_len)
        del k
    ctx.synchronize()
    t_key = time.perf_counter() - t0
    zt = torch.from_numpy(z).pin_memory()
    z_np = zt.numpy().reshape(-1)
    batch = args.batch if args.batch > 0 else world * K
    mine = [i for i in range(batch) if i % world == rank]          # proofs of this rank
    rs = np.random.RandomState(7)
    seeds = rs.randint(1, 1 << 62, size=(batch, 2), dtype=np.int64)  # (r, s) per proof, as StdRng(batch_id) would give

    def prove_range(c, idxs, out):
        for i in idxs:
            out.append(c.prove(pk, m, z_np, int(seeds[i, 0]).to_bytes(32, "little"), int(seeds[i, 1]).to_bytes(32, "little")))

    def step():
        outs = [[] for _ in range(K)]
        thr = [threading.Thread(target=prove_range, args=(ctxs[j], mine[j::K], outs[j])) for j in range(K)]
        for t in thr:
            t.start()
        for t in thr:
            t.join()
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        outs = step()
    if K == 1 and world == 1:
        ctx.profile(True)
        ctx.profile_reset()
    l0 = sum(c.launch_count() for c in ctxs)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outs = step()
    barrier()
    dt = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    phases = {k2: v[0] / args.steps / max(len(mine), 1) for k2, v in ctx.profile_read().items()} if (K == 1 and world == 1) else None
    cpu = None
    if with_cpu:
        from oracle import cpu as orc
        threads = orc.max_threads()
        cpk = orc.ProvingKey(**parts)
        cm = orc.R1cs(ni, nw, csr=(A, B, Cm))
        rb, sb2 = int(seeds[mine[0], 0]).to_bytes(32, "little"), int(seeds[mine[0], 1]).to_bytes(32, "little")
        reps = 3 if lg <= 16 else 1
        t0 = time.perf_counter()
        for _ in range(reps):
            cproof = orc.prove(cpk, cm, z_np, rb, sb2, threads=threads)
        t_cpu = (time.perf_counter() - t0) / reps
        if tuple(bytes(x) for x in ctx.prove(pk, m, z_np, rb, sb2)) != tuple(cproof):
            raise SystemExit("PARITY FAILURE: GPU proof != CPU restatement's proof on the same key, witness and (r, s)")
        cpu = {"value": 1.0 / t_cpu, "unit": "proofs/s", "ms_per_proof": t_cpu * 1e3, "cores": threads, "kind": "port",
               "sample": "%d full proves of the same circuit/key/witness with oracle/cpu_oracle.cpp (arkworks' algorithms restated; "
                         "MSMs parallel over windows only); its proof is byte-identical to the GPU's" % reps}
    launches = (sum(c.launch_count() for c in ctxs) - l0) // max(args.steps, 1)
    if rank == 0:
        line = {"workload": "groth16_prove_synthetic_mimc", "metric": "Groth16 proofs/s (synthetic MiMC circuit, domain 2^%d)" % lg,
                "value": batch / dt, "unit": "proofs/s", "ms_per_step": dt * 1e3, "ms_per_proof_per_context": dt * 1e3 / max(len(mine[0::K]), 1),
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "scaling": "weak" if args.batch == 0 else "strong",
                "config": {"log_domain": lg, "constraints": int(len(A[0]) - 1), "variables": int(nv), "mimc_permutations": num_perm,
                           "proofs_per_step": batch, "contexts_per_gpu": K,
                           "key": ("real Groth16 key: trusted setup run on the GPU (zkb_setup, StdRng(0)) in %.2f s" % t_setup) if with_cpu else
                                  "random curve points generated on the GPU ([k_i]G; no trusted setup: timing-equivalent, proofs do not verify)",
                           "through": "zkb_prove with host z (H2D inside the timed region); wall clock around the batch, max over ranks"},
                "phase_ms_per_proof": phases, "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu,
                "setup_s": {"r1cs_build_host": t_build, "key_generate_and_tables_gpu": t_key},
                "proof_a": bytes(outs[0][0][0]).hex()[:32] if outs[0] else None}
        emit(line)
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


def run_prove_sharded(args):
    """ONE forge-sized proof over N GPUs: key sharded by range, witness map replicated, 768-byte partial records all-gathered
    over NCCL (zelana_b200.multi.ShardedProver).  Strong scaling of the prove latency."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import zelana_b200
    from zelana_b200.multi import ShardedProver
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = zelana_b200.Context(local, stream=stream.cuda_stream)
    lg = args.log_n if args.log_n != 24 else 21
    num_perm = ((1 << lg) - 8) // (4 * 91)
    ni, nw, (A, B, Cm), z = mimc_r1cs_numpy(np, num_perm, seed=0xF0 + lg)
    m = ctx.r1cs(ni, nw, A, B, Cm)
    nv, n = ni + nw, 1 << lg
    k_len = max(nv + 2, n - 1) + 8
    k = rand_fr_range(torch, SEED_BASES, 0, k_len, dev)
    pk = ctx.proving_key_synthetic(nv, nw, n - 1, k, k_len, shard=rank, world=world)
    ctx.synchronize()
    del k
    prover = ShardedProver(ctx, pk, m)
    z_np = torch.from_numpy(z).pin_memory().numpy().reshape(-1)
    r = (123456789).to_bytes(32, "little")
    sb = (987654321).to_bytes(32, "little")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        proof = prover.prove(z_np, r, sb)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proof = prover.prove(z_np, r, sb)
    barrier()
    dt = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        emit({"workload": "groth16_prove_sharded", "metric": "Groth16 prove latency, one proof over N GPUs (synthetic MiMC circuit, domain 2^%d)" % lg,
              "value": dt * 1e3, "unit": "ms", "higher_is_better": False, "scaling": "strong", "n_gpus": world, "steps": args.steps,
              "warmup": args.warmup, "proofs_per_s": 1.0 / dt,
              "config": {"log_domain": lg, "constraints": int(len(A[0]) - 1), "variables": int(nv),
                         "parallelism": "key sharded by range x%d, witness map replicated, all-gather of %d B partial records" % (world, 768),
                         "key": "random curve points generated on the GPU ([k_i]G; timing-equivalent, proofs do not verify)"},
              "clocks": clocks, "proof": (bytes(proof[0]) + bytes(proof[1]) + bytes(proof[2])).hex()[:48]})
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def accumulate_traffic(log_n, world):
    """DRAM bytes per launch of the accumulate kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    capture of THIS round's kernel, profiles/r02_accumulate_traffic.json) for the shard size one rank processes; a number taken
    under a profiler cannot be measured inside this run, so it is the committed capture of the same kernel and size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_accumulate_traffic.json")) as f:
            t = json.load(f)
        shard_log = log_n - (world.bit_length() - 1)
        v = t.get("dram_bytes_per_launch_by_log_points", {}).get(str(shard_log))
        return v if (world & (world - 1)) == 0 else None
    except Exception:
        return None


def measured_hbm_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_sweeps(args):
    """BASELINE.json configs 2 and 3 as sweeps (not the contract line): one JSON line with a row per size."""
    import numpy as np
    import torch
    import zelana_b200
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = zelana_b200.Context(0, stream=stream.cuda_stream)
    peak_int = max(ctx.int32_peak(0)[0], ctx.int32_peak(1)[0])
    hbm, hbm_src = measured_hbm_gbs()
    rows = []

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    if args.workload == "ntt":
        logs = [int(x) for x in args.logs.split(",")] if args.logs else [16, 18, 20, 22, 24, 26]
        for lg in logs:
            n = 1 << lg
            a = rand_fr_range(torch, 0xA77E0000 + lg, 0, n, dev)
            b = torch.empty_like(a)
            row = {"log_n": lg}
            for name, inv, coset in (("ntt", False, False), ("intt", True, False), ("coset_ntt", False, True),
                                     ("coset_intt", True, True)):
                ms = timed(lambda: ctx.ntt_dev(a, b, lg, inverse=inv, coset=coset))
                row[name + "_ms"] = ms
                row[name + "_hbm_frac"] = 64.0 * n / (ms * 1e-3) / (hbm * 1e9)
                row[name + "_int32_frac"] = (n / 2) * lg * 136.0 / (ms * 1e-3) / peak_int
            rows.append(row)
            del a, b
        line = {"workload": "bn254_fr_ntt_sweep", "unit": "ms", "rows": rows, "steps": args.steps, "warmup": args.warmup,
                "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src, "int32_peak_tmul32": peak_int / 1e12,
                "algorithmic_bytes": "64*n per transform", "algorithmic_mul32": "(n/2)*log2(n)*136 per transform",
                "l2": "inputs larger than L2 from 2^22 up; smaller sizes are L2-resident and say so by exceeding the HBM roofline"}
    else:
        logs = [int(x) for x in args.logs.split(",")] if args.logs else [16, 18, 20, 22, 24, 26]
        for lg in logs:
            n = 1 << lg
            k = rand_fr_range(torch, SEED_BASES, 0, n, dev)
            bases = ctx.g1_bases_generate(k, n)
            ctx.synchronize()
            del k
            sc = rand_fr_range(torch, SEED_SCALARS + lg, 0, n, dev)
            out = torch.zeros(64, dtype=torch.uint8, device=dev)
            ms = timed(lambda: ctx.msm_g1_dev(bases, sc, n, out_affine_dev=out))
            row = {"log_n": lg, "ms": ms, "points_per_s": n / (ms * 1e-3),
                   "int32_frac_normalised": MUL32_PER_POINT * n / (ms * 1e-3) / peak_int}
            if lg >= 20:
                # witness-like scalars: 50 % zero, 25 % one, 25 % uniform -- one bucket holds a quarter of the points
                g = torch.Generator(device=dev)
                g.manual_seed(0x717 + lg)
                u = torch.rand(n, device=dev, generator=g)
                wl = sc.clone()
                wl[u < 0.5] = 0
                one = torch.zeros(8, dtype=torch.int32, device=dev)
                one[0] = 1
                wl[(u >= 0.5) & (u < 0.75)] = one
                row["witness_like_ms"] = timed(lambda: ctx.msm_g1_dev(bases, wl, n, out_affine_dev=out))
                del wl, u
            rows.append(row)
            bases.free()
            del sc
        line = {"workload": "bn254_g1_msm_sweep", "unit": "ms", "rows": rows, "steps": args.steps, "warmup": args.warmup,
                "int32_peak_tmul32": peak_int / 1e12, "normalisation": "21760 mul32 per point (SURVEY 8d)"}
    emit(line)
    ctx.close()


class _CleanStdout:
    """Everything any library prints to fd 1 (NCCL prints its version there) goes to stderr; only our JSON line reaches
    the real stdout, so the driver always finds exactly one line to parse."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        self.out = os.fdopen(self.real, "w", closefd=False)
        return self

    def emit(self, obj):
        self.out.write(json.dumps(obj) + "\n")
        self.out.flush()

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


_OUT = None


def emit(obj):
    if _OUT is not None:
        _OUT.emit(obj)
    else:
        print(json.dumps(obj), flush=True)


def main():
    global _OUT
    args = parse_args()
    with _CleanStdout() as _OUT:
        _main(args)
    _OUT = None


def _main(args):
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "prove":
        run_prove(args)
    elif args.workload == "prove_sharded":
        run_prove_sharded(args)
    elif args.workload != "msm_g1":
        run_sweeps(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
