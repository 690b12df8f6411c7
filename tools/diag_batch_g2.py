import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zelana_b200
from tools.diag_batch import rand_fr

ctx = zelana_b200.Context(0)
for group in (1, 2):
    for n, K in ((300, 2), (300, 3), (3000, 9)):
        k = rand_fr(n, 81)
        gen = ctx.g1_bases_generate if group == 1 else ctx.g2_bases_generate
        msm = ctx.msm_g1 if group == 1 else ctx.msm_g2
        bases = gen(torch.from_numpy(k.view(np.int32)).cuda(), n)
        c, nwin = bases.window()
        sc = rand_fr(n * K, 83).reshape(K, n, 8)
        sd = torch.from_numpy(sc.view(np.int32).copy()).cuda()
        single = [msm(bases, sc[p]) for p in range(K)]
        gotb = ctx.debug_msm_batch(group, bases, sd, n, n, K)
        print("group %d n=%d K=%d c=%d: random vectors" % (group, n, K, c), [a == b for a, b in zip(gotb, single)])
        sp = sc.copy()
        sp[1] = 0
        if K > 2:
            sp[2, ::2] = 0
        if K > 3:
            sp[3, :, 1:] = 0
            sp[3, :, 0] = 1
        single = [msm(bases, sp[p]) for p in range(K)]
        gotb = ctx.debug_msm_batch(group, bases, torch.from_numpy(sp.view(np.int32).copy()).cuda(), n, n, K)
        print("      special vectors (1: zeros, 2: half zeros, 3: ones)", [a == b for a, b in zip(gotb, single)])
        one = sc.copy()
        one[:, :, 1:] = 0
        one[:, :, 0] = 1
        single = [msm(bases, one[p]) for p in range(K)]
        gotb = ctx.debug_msm_batch(group, bases, torch.from_numpy(one.view(np.int32).copy()).cuda(), n, n, K)
        print("      all ones", [a == b for a, b in zip(gotb, single)])
        bases.free()
ctx.close()
