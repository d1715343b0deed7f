"""Diagnostics: per-stream cost (SM cycles) vs simple host-side features, contiguous shards.  Writes gpurun_out/streams.npz"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trust4_b200 import api, synth
pairs, S = int(sys.argv[1]), int(sys.argv[2])
lib = api.default_lib(); lib.check(lib.init(0, 0))
cl = synth.make_clones(max(20, pairs // 50), 1)
rd = synth.sample_pairs(cl, pairs, 150, 1000)
w = synth.build_workload(cl, rd, device=torch.device("cuda", 0))
off, descs = synth.shard_workload(w, S, deal=False)
sets = api.SeqSet.create_many(S, 9, lib)
ret, st, resc = api.streams_run(sets, synth.run_cfg(), descs, off, w.pool, w.names, lib)
hs = (C.c_void_p * S)(*[s.h for s in sets])
cyc = np.zeros(S, dtype=np.uint64); lib.check(lib.streams_cycles(hs, S, cyc.ctypes.data))
nd = np.add.reduceat(((descs["flags"] & 1) == 0).astype(np.int64), off[:-1])
sizes = np.array([s.size() for s in sets])
np.savez("gpurun_out/streams.npz", cyc=cyc, nd=nd, sizes=sizes, off=off, mincnt=np.add.reduceat(descs["min_cnt"].astype(np.int64), off[:-1]))
print("contigs total", sizes.sum(), "reads", len(descs), "assembled", int((ret >= 0).sum() + (resc >= 0).sum()))
