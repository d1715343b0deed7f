#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Dev-container only: needs /root/reference (sources + example data) and
oracle/_ref (run `make -C oracle` first).  For each case it runs the stock
stage-1 binary and the call-tracing build, checks that both write identical
_raw.out / _final.out, and stores
    <case>.trace.gz   every SeqSet call of the stage-1 driver with its arguments
                      and return values (format: tests/trace_format.md)
    <case>_raw.out.gz the reference's own stage-1 output (SeqSet::Output)
Cases: config 1 (shipped example, BASELINE.json configs[0]) and two small seeded
synthetic sets through the reference's full pre-processing.
"""
import gzip, os, shutil, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(name, args, tmp):
    env = dict(os.environ)
    subprocess.run([os.path.join(BIN, "trust4"), "-t", "1", "-o", os.path.join(tmp, name + "_stock")] + args,
                   check=True, stderr=subprocess.DEVNULL)
    env["T4_TRACE_OUT"] = os.path.join(tmp, name + ".trace")
    subprocess.run([os.path.join(BIN, "trust4_trace"), "-t", "1", "-o", os.path.join(tmp, name + "_trace")] + args,
                   check=True, env=env, stderr=subprocess.DEVNULL)
    for suf in ("_raw.out", "_final.out", "_assembled_reads.fa"):
        a = open(os.path.join(tmp, name + "_stock" + suf), "rb").read()
        b = open(os.path.join(tmp, name + "_trace" + suf), "rb").read()
        assert a == b, (name, suf)
    for src, dst in ((name + ".trace", name + ".trace.gz"), (name + "_stock_raw.out", name + "_raw.out.gz")):
        with open(os.path.join(tmp, src), "rb") as f, gzip.GzipFile(os.path.join(OUT, dst), "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
    print(name, "ok")


def main():
    from trust4_b200 import synth
    with tempfile.TemporaryDirectory() as tmp:
        run_case("example", ["-f", REF + "/hg38_bcrtcr.fa", "-1", REF + "/example/example_1.fq", "-2", REF + "/example/example_2.fq"], tmp)
        for name, npairs, nclones, seed in (("synth2k", 2000, 40, 11), ("synth6k", 6000, 300, 12)):
            cl = synth.make_clones(nclones, seed)
            rd = synth.sample_pairs(cl, npairs, 150, seed)
            synth.write_fastq(rd, os.path.join(tmp, name))
            run_case(name, ["-f", REF + "/human_IMGT+C.fa", "-1", os.path.join(tmp, name + "_1.fq"), "-2", os.path.join(tmp, name + "_2.fq")], tmp)


if __name__ == "__main__":
    main()
