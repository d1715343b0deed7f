#!/usr/bin/env python3
"""End-to-end stage-1 CLI comparison on the GPU box: stock `trust4 -t <cores>` (oracle/_ref/trust4, the reference compiled
from its own sources) vs the drop-in built from the same main.cpp (integration/_build/trust4_gpu_batch), same FASTQ.

    python bench/cli_compare.py [--pairs 50000] [--streams 1,64,1024] [--shard-by gene,rank] > gpurun_out/cli_compare.json

Reports, per run, the contiguity of the `_raw.out` / `_final.out` contigs against the synthetic truth (bench/quality.py:
clonotypes whose V(D)J core lies in ONE contig), the wall time of the whole binary and of the AddRead loop (from the reference's own log lines "Finish rough
annotations." -> "Assembled %d reads.", main.cpp:1121, 1883), and whether the outputs are byte-identical (S = 1 must be)."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bench"))
from trust4_b200 import synth  # noqa: E402
import quality  # noqa: E402

STAMP = re.compile(r"^\[(\w+ \w+\s+\d+ \d+:\d+:\d+ \d+)\] (.*)$")


def loop_seconds(log):
    t = {}
    for ln in log.splitlines():
        m = STAMP.match(ln.strip())
        if not m:
            continue
        ts = time.mktime(time.strptime(m.group(1), "%a %b %d %H:%M:%S %Y"))
        if m.group(2).startswith("Finish rough annotations"):
            t["a"] = ts
        if m.group(2).startswith("Assembled"):
            t["b"] = ts
    return (t["b"] - t["a"]) if "a" in t and "b" in t else None


def run(exe, args, out, env=None):
    t0 = time.perf_counter()
    p = subprocess.run([exe] + args + ["-o", out], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env, timeout=3600)
    wall = time.perf_counter() - t0
    log = p.stderr.decode(errors="replace")
    return {"rc": p.returncode, "wall_s": wall, "addread_loop_s_from_log": loop_seconds(log),
            "assign_pass_on_device": "AssignRead pass on the device" in log}


def contiguity(cl, rd, prefix):
    out = {}
    for suf in ("_raw.out", "_final.out"):
        try:
            codes, off, n = quality.contigs_from_output(open(prefix + suf, "rb").read())
            q = quality.spanning_fraction(cl, rd, codes, off)
            out[suf] = {"contigs": n, "spanned_fraction": q["spanned_fraction"], "covered_by_reads": q["covered_by_reads"]}
        except Exception as e:       # noqa: BLE001
            out[suf] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=50000)
    ap.add_argument("--streams", default="1,64,1024")
    ap.add_argument("--shard-by", default="gene,rank", help="T4_SHARD_BY values tried for S > 1")
    ap.add_argument("--batch-binary", default=os.path.join(ROOT, "integration", "_build", "trust4_gpu_batch"))
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="t4cli")
    pool = synth.load_gene_pool()
    fa = os.path.join(tmp, "genes.fa")
    with open(fa, "w") as f:
        for ch in pool.values():
            for seg in ch.values():
                for name, seq in seg:
                    f.write(">%s\n%s\n" % (name, seq))
    cl = synth.make_clones(max(20, a.pairs // 50), 1)
    rd = synth.sample_pairs(cl, a.pairs, 150, 1)
    synth.write_fastq(rd, os.path.join(tmp, "r"))
    cores = os.cpu_count() or 1
    args = ["-f", fa, "-1", os.path.join(tmp, "r_1.fq"), "-2", os.path.join(tmp, "r_2.fq"), "-t", str(cores)]
    stock = os.path.join(ROOT, "oracle", "_ref", "trust4")
    batch = a.batch_binary
    res = {"pairs": a.pairs, "host_threads": cores, "runs": []}
    r = run(stock, args, os.path.join(tmp, "stock"))
    r["binary"] = "stock trust4 -t %d" % cores
    r["contiguity"] = contiguity(cl, rd, os.path.join(tmp, "stock"))
    res["runs"].append(r)
    combos = [(S, by) for S in [int(x) for x in a.streams.split(",")] for by in (a.shard_by.split(",") if S > 1 else ["gene"])]
    for S, by in combos:
        tag = "s%d%s" % (S, by)
        r = run(batch, args, os.path.join(tmp, tag), env=dict(os.environ, T4_STREAMS=str(S), T4_SHARD_BY=by))
        r["binary"] = "trust4_gpu_batch T4_STREAMS=%d%s -t %d" % (S, " T4_SHARD_BY=" + by if S > 1 else "", cores)
        r["contiguity"] = contiguity(cl, rd, os.path.join(tmp, tag))
        same = {}
        for suf in ("_raw.out", "_final.out", "_assembled_reads.fa"):
            try:
                same[suf] = open(os.path.join(tmp, "stock" + suf), "rb").read() == open(os.path.join(tmp, tag + suf), "rb").read()
            except Exception:
                same[suf] = None
        r["identical_to_stock"] = same
        try:
            r["raw_contigs"] = open(os.path.join(tmp, tag + "_raw.out"), "rb").read().count(b">")
        except Exception:
            pass
        res["runs"].append(r)
    try:
        res["stock_raw_contigs"] = open(os.path.join(tmp, "stock_raw.out"), "rb").read().count(b">")
    except Exception:
        pass
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
