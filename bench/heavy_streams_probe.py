"""Diagnostics: phase shares of the heaviest contiguous shards only (the critical path of a contiguous-shard launch)."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trust4_b200 import api, synth
pairs, S, lo, hi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lib = api.default_lib(); lib.check(lib.init(0, 0))
cl = synth.make_clones(max(20, pairs // 50), 1)
rd = synth.sample_pairs(cl, pairs, 150, 1000)
w = synth.build_workload(cl, rd, device=torch.device("cuda", 0))
off, descs = synth.shard_workload(w, S, deal=False)
sub_off = off[lo:hi + 1] - off[lo]
sub = descs[off[lo]:off[hi]].copy()
n = hi - lo
sets = api.SeqSet.create_many(n, 9, lib)
lib.check(lib.reset()); sets = api.SeqSet.create_many(n, 9, lib)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ret, st, resc = api.streams_run(sets, synth.run_cfg(), sub, sub_off, w.pool, w.names, lib)
e1.record(); torch.cuda.synchronize()
c = np.zeros(api.N_COUNTERS, dtype=np.uint64); lib.check(lib.last_counters(c.ctypes.data)); c = c.astype(float)
names = ["other", "probe", "hit_sort", "chains", "score", "extend", "decide_commit", "novel_repeat_consensus"]
nr = len(sub)
print("ms", e0.elapsed_time(e1), "reads", nr, dict(zip(names, (c[8:16] / c[8:16].sum()).round(3))),
      {"hits": c[4] / nr, "ovl_scored": c[6] / nr, "ovl_ext": c[16] / nr, "ext_dps": c[1] / nr, "gap_dps": c[7] / nr,
       "ext_bits_cyc": c[17] / max(1, c[6] and nr), "ext_classify_cyc": c[18] / nr, "ext_dp_cyc": c[19] / nr, "easy_frac": c[20] / nr})
