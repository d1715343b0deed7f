# One-off check (CPU, ~9 min): the exact bench.py workload (1 M pairs, 4096 contiguous shards, seeds as bench.py rank 0)
# through the engine emulation vs the compiled reference, every shard: return codes, strands, rescue codes, Output
# text and index checksum.  Result 2026-09-23: 4096 of 4096 shards identical, 1 999 998 reads assembled, 425 202 contigs.
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from trust4_b200 import api, synth
import refharness as rh
from concurrent.futures import ThreadPoolExecutor
lib = api.Lib('/root/repo/tests/emu/libt4emu.so', 't4emu_'); lib.check(lib.init(0, 40<<30))
pairs, S = int(sys.argv[1]), int(sys.argv[2])
t=time.time()
cl = synth.make_clones(max(20, pairs//50), 1); rd = synth.sample_pairs(cl, pairs, 150, 1000); w = synth.build_workload(cl, rd)
off, d = synth.shard_workload(w, S)
print('workload', time.time()-t, flush=True)
cfg = synth.run_cfg()
t=time.time(); sets = api.SeqSet.create_many(S, 9, lib); ret, st, resc = api.streams_run(sets, cfg, d, off, w.pool, w.names, lib); print('emu', time.time()-t, 'assembled', int((ret>=0).sum()+(resc>=0).sum()), flush=True)
outs = [s.output() for s in sets]; idx = [s.index_checksum() for s in sets]
def check(j):
    r = rh.RefSeqSet(9); lo,hi=int(off[j]),int(off[j+1])
    a, rret, rstr, rresc = r.run_descs(cfg, d[lo:hi].copy(), w.pool, w.names)
    ok = (rret==ret[lo:hi]).all() and (rstr==st[lo:hi]).all() and (rresc==resc[lo:hi]).all() and r.output()==outs[j] and r.index_checksum()==idx[j]
    n = r.size(); r.close(); return ok, n
t=time.time()
with ThreadPoolExecutor(8) as ex: res = list(ex.map(check, range(S)))
print('ref', time.time()-t, 'shards ok', sum(1 for o,_ in res if o), 'of', S, 'contigs', sum(n for _,n in res), flush=True)
