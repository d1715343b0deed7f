"""bench.py's JSON line: the keys the driver reads must survive edits (round 1 lost roofline / cpu_baseline to a trailing
comment), and the reference sample must visit shards in a fixed, uniform order."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_build_line_has_every_contract_key():
    args = bench.parse([])
    roof = {"bound": "hbm", "achieved": 1.0, "peak": 2.0, "unit": "GB/s", "frac": 0.5, "traffic": None}
    cpu = {"value": 1.0, "unit": "reads/s", "cores": 1, "kind": "reference", "sample": "x"}
    line = bench.build_line(args, {"workload": "w"}, 1.0, 2.0, {"sm_mhz": 1.0, "sm_max_mhz": 2.0, "reasons": []},
                            {"value": 1.0, "unit": "reads/s", "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 1}, 4, roof, dict(roof), cpu,
                            {"assembled_reads": 1})
    s = json.dumps(line)
    back = json.loads(s)
    for k in bench.REQUIRED_KEYS:
        assert k in back, k
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in back["roofline"] and k in back["roofline_probe"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in back["cpu_baseline"]
    assert back["higher_is_better"] is True and back["n_gpus"] == 1


def test_bench_source_has_no_key_hidden_in_a_comment():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for ln in src.splitlines():
        code, _, comment = ln.partition("#")
        assert '"roofline":' not in comment and '"cpu_baseline":' not in comment, ln


def test_reference_sample_order_is_a_fixed_uniform_permutation():
    o = bench.sample_order(4096)
    assert sorted(o.tolist()) == list(range(4096))
    assert (o == bench.sample_order(4096)).all()
    # any prefix covers the index range evenly: mean index of the first 256 within 10 % of the centre
    assert abs(np.mean(o[:256]) - 2048) < 205


def test_bench_source_emits_the_separate_figures():
    """The AssignRead pass and the k-mer statistics are separate figures of the line (outside value / e2e): their keys are
    passed to build_line, are not hidden in a comment, and both legs are guarded so that they cannot break the headline."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"assign_pass": assign_fig', '"preprocess_kmer_stats": kc_fig'):
        assert any(key in ln.partition("#")[0] for ln in src.splitlines()), key
    assert src.count("except Exception as ex:") >= 4
    args = bench.parse([])
    assert args.assign_pass == 1 and args.kmer_stats == 1
