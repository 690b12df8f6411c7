"""Where do small (L2 batch circuit, domain 2^13) proofs spend their time?  Sweeps the number of contexts per GPU with the
CUDA-graph replay on and off, with and without the host-side witness assignment, and times one proof alone.

    python tools/l2_throughput.py [--contexts 1,2,4,8,12,16] [--per-context 16] > gpurun_out/l2_throughput.json
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contexts", default="1,2,4,8,12,16")
    ap.add_argument("--per-context", type=int, default=16)
    args = ap.parse_args()
    import torch
    import zelana_b200
    from zelana_b200 import l2_circuit as l2
    ks = [int(x) for x in args.contexts.split(",")]
    kmax = max(ks)
    dev = torch.device("cuda", 0)
    streams = [torch.cuda.Stream(device=dev) for _ in range(kmax)]
    ctxs = [zelana_b200.Context(0, stream=st.cuda_stream) for st in streams]
    circ, pk_bytes, vk_bytes, _raw = l2.keygen(ctxs[0])
    pk = ctxs[0].proving_key_compressed(pk_bytes, validate=False)
    a, b, c = circ.matrices()
    m = ctxs[0].r1cs(circ.num_instance, circ.num_witness, a, b, c)
    lib = circ.lib

    def batch(bid):
        ckt = l2.L2BlockCircuit(transactions=[l2.TransactionWitness(bytes([1] * 32), bytes([2] * 32), 1 + bid % 1000)],
                                initial_accounts={bytes([1] * 32): 1000, bytes([2] * 32): bid}, batch_id=bid)
        ckt = ckt.with_inputs(l2.satisfying_inputs(ckt))
        w, keep = l2._c_witness(ckt)
        z = circ.assign(ckt)
        r, s = l2.prover_randomness(bid)
        return ckt, w, keep, l2._c_inputs(ckt.public_inputs()), (C.c_uint8 * 256)(), z, r, s

    jobs = [[batch(j * args.per_context + i + 1) for i in range(args.per_context)] for j in range(kmax)]
    oa, ob, oc = (C.c_uint8 * 64)(), (C.c_uint8 * 128)(), (C.c_uint8 * 64)()

    def work_full(cx, mine):
        for ckt, w, keep, x, out, z, r, s in mine:
            rc = lib.zkb_l2_prove(cx.h, pk.h, m.h, circ.h, C.byref(x), C.byref(w), out)
            assert rc == 0, lib.zkb_l2_last_error()

    def work_gpu_only(cx, mine):
        a_, b_, c_ = (C.c_uint8 * 64)(), (C.c_uint8 * 128)(), (C.c_uint8 * 64)()
        for ckt, w, keep, x, out, z, r, s in mine:
            rc = lib.zkb_prove(cx.h, pk.h, m.h, z, r, s, a_, b_, c_)
            assert rc == 0

    def run(K, fn):
        th = [threading.Thread(target=fn, args=(ctxs[j], jobs[j])) for j in range(K)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    rows = []
    for graphs in (False, True):
        for cx in ctxs:
            cx.set_graphs(graphs)
        for K in ks:
            for name, fn in (("l2_prove (assign + prove)", work_full), ("prove only (z ready)", work_gpu_only)):
                run(K, fn)
                dt = min(run(K, fn) for _ in range(3))
                rows.append({"graphs": graphs, "contexts": K, "what": name, "proofs_per_s": K * args.per_context / dt,
                             "ms_per_proof_per_context": dt / args.per_context * 1e3})
    # one proof alone: wall latency of zkb_prove and of the assignment
    cx = ctxs[0]
    lat = {}
    for graphs in (False, True):
        cx.set_graphs(graphs)
        work_gpu_only(cx, jobs[0][:3])
        t0 = time.perf_counter()
        work_gpu_only(cx, jobs[0])
        lat["prove_ms_graphs_%s" % ("on" if graphs else "off")] = (time.perf_counter() - t0) / args.per_context * 1e3
    t0 = time.perf_counter()
    for j in jobs[0]:
        circ.assign(j[0])
    lat["assign_ms"] = (time.perf_counter() - t0) / args.per_context * 1e3
    lat["graph_stats_ctx0"] = cx.graph_stats()
    lat["launches_per_proof"] = None
    l0 = cx.launch_count()
    work_gpu_only(cx, jobs[0][:1])
    lat["launches_per_proof"] = cx.launch_count() - l0
    print(json.dumps({"rows": rows, "single": lat, "host_cores": os.cpu_count()}, indent=1))


if __name__ == "__main__":
    main()
