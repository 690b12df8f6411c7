import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zelana_b200
from tools.diag_batch import rand_fr

ctx = zelana_b200.Context(0)
n, K = 300, 3
k = rand_fr(n, 81)
sc = rand_fr(n * K, 83).reshape(K, n, 8)
sd = torch.from_numpy(sc.view(np.int32).copy()).cuda()
for group in (2, 1):
    for c in (4, 6, 7, 8, 10, 11, 12, 13):
        ctx.set_msm_window(c)
        gen = ctx.g1_bases_generate if group == 1 else ctx.g2_bases_generate
        msm = ctx.msm_g1 if group == 1 else ctx.msm_g2
        bases = gen(torch.from_numpy(k.view(np.int32)).cuda(), n)
        single = [msm(bases, sc[p]) for p in range(K)]
        gotb = ctx.debug_msm_batch(group, bases, sd, n, n, K)
        print("group %d c=%d nsp=%d" % (group, c, (1 << (c - 1)) >> min(c - 1, 5)), [a == b for a, b in zip(gotb, single)])
        bases.free()
ctx.close()
