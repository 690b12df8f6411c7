"""G2 vs G1 MSM of the same size and scalars, per-phase CUDA-event times (zkb_prof_*):  python tools/prof_g2.py [log_n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import zelana_b200  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
dev = torch.device("cuda", 0)
ctx = zelana_b200.Context(0)
g = torch.Generator(device=dev)
g.manual_seed(1)


def rnd():
    k = torch.randint(0, 1 << 32, (n, 8), dtype=torch.int64, device=dev, generator=g)
    k[:, 7] %= 0x30644E72
    return k.to(torch.int32)


k, s = rnd(), rnd()
for group in (1, 2):
    bases = (ctx.g1_bases_generate if group == 1 else ctx.g2_bases_generate)(k, n)
    out = torch.zeros(64 * group, dtype=torch.uint8, device=dev)
    fn = ctx.msm_g1_dev if group == 1 else ctx.msm_g2_dev
    for _ in range(2):
        fn(bases, s, n, out_affine_dev=out)
    torch.cuda.synchronize()
    ctx.profile(True)
    ctx.profile_reset()
    for _ in range(3):
        fn(bases, s, n, out_affine_dev=out)
    torch.cuda.synchronize()
    print("G%d 2^%d:" % (group, log_n), {k2: round(v[0] / 3, 3) for k2, v in ctx.profile_read().items() if v[0] > 0}, "window", bases.window())
    ctx.profile(False)
    bases.free()
