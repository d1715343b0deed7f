#!/usr/bin/env python3
"""Stage-0 candidate extraction figure: fastq-extractor's per-read predicate on the device.

    python bench/stage0_scan.py [--reads 2000000] > gpurun_out/stage0_scan.json

Half of the reads are 150 bp reads of synthetic clonotypes (candidates), half are random 150-mers (what a real RNA-seq
library mostly consists of).  Reports reads/s of t4_refset_scan_device (CUDA events, inputs resident), the parity of a sample
against the reference's IsLowComplexity / HasHitInSet(read, 0), and the reference's own rate on the host threads."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2000000)
    ap.add_argument("--sample", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import torch
    from trust4_b200 import api, synth
    import parity_cases as pc
    import refharness as rh
    lib = api.default_lib()
    lib.check(lib.init(0, 0))
    dev = torch.device("cuda", 0)
    tmp = tempfile.mkdtemp(prefix="t4s0")
    fa = os.path.join(tmp, "genes.fa")
    pc.write_gene_fasta(fa, messy=False)
    L = 150
    n_c = a.reads // 2
    cl = synth.make_clones(max(20, n_c // 100), a.seed)
    rd = synth.sample_pairs(cl, n_c // 2, L, a.seed)
    rng = np.random.default_rng(a.seed)
    codes = np.concatenate([rd.codes, rng.integers(0, 4, size=(a.reads - len(rd.codes), L), dtype=np.uint8)])
    perm = rng.permutation(len(codes))
    codes = codes[perm]
    n = len(codes)
    pool = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].reshape(-1)
    hit_len = max(27, L // 5)                     # FastqExtractor.cpp:436-455
    g = api.RefSet(fa, 9, lib, hit_len_required=hit_len)
    dpool = torch.from_numpy(np.concatenate([pool, np.zeros(64, dtype=np.uint8)])).to(dev)
    doff = torch.arange(n, dtype=torch.int64, device=dev) * L
    dlen = torch.full((n,), L, dtype=torch.int32, device=dev)
    ostr = torch.zeros(n, dtype=torch.int8, device=dev)
    olow = torch.zeros(n, dtype=torch.uint8, device=dev)
    ctrl = torch.zeros(8, dtype=torch.int64, device=dev)
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        lib.check(lib.refset_scan_device(g.h, dpool.data_ptr(), doff.data_ptr(), dlen.data_ptr(), n, ostr.data_ptr(), olow.data_ptr(),
                                         ctrl.data_ptr(), 0, None))
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    gs = ostr.cpu().numpy()
    gl = olow.cpu().numpy()
    c = ctrl.cpu().numpy()
    out = {"what": "fastq-extractor predicate on the device: IsLowComplexity + HasHitInSet(read, 0) against the gene set (t4_refset_scan_device, "
                   "worker CTAs of t4_aux_kernel)", "reads": n, "read_len": L, "candidate_fraction_generated": float(len(rd.codes)) / n,
           "genes": g.size(), "hit_len_required": hit_len, "ms_all": ms, "ms": float(np.median(ms)), "reads_per_s": n / (np.median(ms) * 1e-3),
           "with_hit": int(c[1]), "low_complexity": int(c[2]), "kept_by_extractor": int(((gs != 0) & (gl == 0)).sum())}
    if rh.available():
        r = rh.RefGeneSet(fa, 9, hit_len_required=hit_len)
        idx = rng.choice(n, size=min(a.sample, n), replace=False)
        reads = [bytes(pool[i * L:(i + 1) * L]).decode() for i in idx]
        cores = os.cpu_count() or 1
        res = [None] * len(reads)

        def work(t):
            for j in range(t, len(reads), cores):
                res[j] = (r.has_hit_in_set(reads[j], 0), rh.is_low_complexity(reads[j]))

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        el = time.perf_counter() - t0
        rs = np.array([x[0] for x in res], dtype=np.int8)
        rl = np.array([x[1] for x in res], dtype=np.uint8)
        out["parity_sample"] = {"reads": len(reads), "equal_reference": bool((gs[idx] == rs).all() and (gl[idx] == rl).all()),
                                "mismatches": int((gs[idx] != rs).sum() + (gl[idx] != rl).sum())}
        out["cpu_reference"] = {"reads_per_s": len(reads) / el, "threads": cores, "seconds": el,
                                "note": "reference SeqSet::HasHitInSet + IsLowComplexity over the sample, ctypes calls from Python threads (GIL released inside the call)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
