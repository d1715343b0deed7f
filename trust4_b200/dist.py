"""Merge-step exchange for read-sharded multi-GPU runs: one all-gather of the per-rank packed
contig buffers (t4_streams_pack_contigs).  Works with NCCL (device tensors) and gloo (CPU tensors, tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np


def pack_contigs(lib, sets, device=None):
    """Returns a uint8 torch tensor (on `device`, or CPU for the emulation) with every live contig of `sets`."""
    import torch
    hs = (C.c_void_p * len(sets))(*[s.h for s in sets])
    need, n = C.c_size_t(), C.c_int64()
    lib.check(lib.streams_pack_contigs(hs, len(sets), None, 0, C.byref(need), C.byref(n)))
    buf = torch.empty(max(16, need.value), dtype=torch.uint8, device=device or "cpu")
    lib.check(lib.streams_pack_contigs(hs, len(sets), buf.data_ptr(), buf.numel(), C.byref(need), C.byref(n)))
    return buf[: need.value], int(n.value)


def allgather_contigs(buf):
    """All ranks receive the per-rank packed buffers, in rank order: the sizes first (one small all-gather), then ONE
    all_gather_into_tensor on a flat buffer padded to the largest size (NCCL: a single collective, no per-rank copies)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([buf.numel()], dtype=torch.int64, device=buf.device)
    sizes_t = torch.zeros(world, dtype=torch.int64, device=buf.device)
    dist.all_gather_into_tensor(sizes_t, n)
    sizes = [int(x) for x in sizes_t.tolist()]
    mx = max(max(sizes), 16)
    mx = (mx + 15) & ~15
    base = buf._base if getattr(buf, "_base", None) is not None else buf
    if base.numel() >= mx and buf.storage_offset() == 0 and base.dim() == 1:
        padded = base[:mx]                 # the pack buffer has slack: bytes past this rank's size are never read back
    else:
        padded = torch.zeros(mx, dtype=torch.uint8, device=buf.device)
        padded[: buf.numel()] = buf
    out = torch.empty(world * mx, dtype=torch.uint8, device=buf.device)
    dist.all_gather_into_tensor(out, padded)
    return [out[r * mx: r * mx + sizes[r]] for r in range(world)]


def unpack_contigs(buf):
    """Host-side decode of a packed buffer -> list of dicts (set, slot, consensus, pos_weight[len,4], name, barcode, num_read)."""
    a = buf.cpu().numpy() if hasattr(buf, "cpu") else np.asarray(buf)
    out = []
    o = 0
    while o < len(a):
        h = a[o:o + 32].view(np.uint32)
        st, slot, ln, nl, rb = int(h[0]), int(h[1]), int(h[2]), int(h[3]), int(h[6])
        bc, nr = int(h[4:6].view(np.int32)[0]), int(h[4:6].view(np.int32)[1])
        cons = a[o + 32:o + 32 + ln].tobytes().decode()
        if int(h[7]) & 1:       # posWeight counts stored as u16
            pw = a[o + 32 + ln:o + 32 + 9 * ln].copy().view(np.uint16).reshape(ln, 4).astype(np.int32)
            na = o + 32 + 9 * ln
        else:
            pw = a[o + 32 + ln:o + 32 + 17 * ln].copy().view(np.int32).reshape(ln, 4)
            na = o + 32 + 17 * ln
        name = a[na:na + nl].tobytes().decode()
        out.append(dict(set=st, slot=slot, consensus=cons, pos_weight=pw, name=name, barcode=bc, num_read=nr))
        o += rb
    return out


def format_output(contigs):
    """SeqSet::Output text (SeqSet.hpp:10939) of unpacked contigs, per set."""
    chunks = {}
    for c in contigs:
        s = [">assemble%d %s\n%s\n" % (c["slot"], c["name"], c["consensus"])]
        for k in range(4):
            s.append("".join("%d " % v for v in c["pos_weight"][:, k]) + "\n")
        chunks.setdefault(c["set"], []).append("".join(s))
    return {k: "".join(v).encode() for k, v in chunks.items()}
