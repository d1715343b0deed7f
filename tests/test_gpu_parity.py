"""Parity tests proper: the CUDA engine through the C ABI (libtrust4_b200.so) against the oracle --
the reference's own classes compiled into oracle/_ref/libt4ref.so (travels to the GPU box prebuilt) and the
golden call traces / outputs of the stock trust4 binary in tests/golden/."""
import numpy as np
import pytest

import parity_cases as pc
from trust4_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["example", "synth2k", "synth6k"])
def test_gpu_trace_replay(gpu_lib, name):
    """BASELINE.json configs[0] (shipped example) and two synthetic sets: every AddRead / RepeatAddRead /
    InputNovelRead / UpdateAllConsensus call of the real stage-1 run, bit-exact returns and _raw.out."""
    pc.check_trace_replay(gpu_lib, name)


@pytest.mark.parametrize("seed,shards", [(1, 1), (2, 3), (3, 8), (4, 64)])
def test_gpu_batch_vs_reference(gpu_lib, ref, seed, shards):
    assert pc.check_batch_vs_ref(gpu_lib, ref, seed, shards, nclones=40, npairs=1500) > 100


@pytest.mark.parametrize("seed,shards", [(11, 7), (12, 96)])
def test_gpu_batch_dealt_shards(gpu_lib, ref, seed, shards):
    """Runs of identical reads dealt round-robin to the streams (bench.py's default sharding)."""
    assert pc.check_batch_vs_ref(gpu_lib, ref, seed, shards, nclones=60, npairs=2500, deal=True) > 100


def test_gpu_batch_larger_single_stream(gpu_lib, ref):
    """One stream, enough contigs for the periodic UpdateAllConsensus (every 10000 assembled reads)."""
    assert pc.check_batch_vs_ref(gpu_lib, ref, 9, 1, nclones=300, npairs=6000) > 10000


def test_gpu_stage_parity(gpu_lib, ref):
    """k-mer hits, chains -> scored overlaps (similarity doubles bit-equal), index multiset."""
    pc.check_stage_parity(gpu_lib, ref, "synth2k", every=37, max_checks=80)


def test_gpu_dp(gpu_lib, ref):
    for seed in (5, 6, 7):
        pc.check_dp(gpu_lib, ref, seed)


def test_gpu_dp_hot_path_register(gpu_lib, ref):
    """t4_dp_equal -- the per-thread register banded DP ExtendOverlap runs -- score + edit string vs the reference."""
    for seed in (11, 12):
        pc.check_dp_hot(gpu_lib, ref, seed, 0)


def test_gpu_dp_hot_path_half_warp(gpu_lib, ref):
    """w_dp_equal_half -- the half-warp anti-diagonal DP of gap scoring -- score + edit string vs the reference."""
    for seed in (11, 12):
        pc.check_dp_hot(gpu_lib, ref, seed, 1)


def test_gpu_change_kmer_length(gpu_lib, ref):
    """ChangeKmerLength (slot compaction + full re-index) in the middle of a stream."""
    gpu_lib.check(gpu_lib.reset())
    w = pc.small_workload(21, 60, 1500)
    cfg = synth.run_cfg(change_k_threshold=20)     # force k: 9 -> 11 -> 13 early
    g = api.SeqSet(9, gpu_lib)
    r = ref.RefSeqSet(9)
    _, gret, gstr, gres = g.run_descs(cfg, w.descs, w.pool, w.names)
    _, rret, rstr, rres = r.run_descs(cfg, w.descs, w.pool, w.names)
    assert g.kmer_length() == r.kmer_length() and g.kmer_length() > 9
    assert (gret == rret).all() and (gstr == rstr).all() and (gres == rres).all()
    assert g.output() == r.output()
    assert g.index_checksum() == r.index_checksum()


def test_gpu_empty_and_short_inputs(gpu_lib, ref):
    gpu_lib.check(gpu_lib.reset())
    g = api.SeqSet(9, gpu_lib)
    r = ref.RefSeqSet(9)
    assert g.output() == r.output() == b""
    for read in ("ACGT", "ACGTACGTA", "A" * 60, "ACGTNACGTACGTTTGACCANNACGATCGATCGATTTACGACGGGATCTAGCAGGACTTT"):
        assert g.add_read(read, "", 0, -1, 1, 0, 0.9) == r.add_read(read, "", 0, -1, 1, 0, 0.9)
        assert g.repeat_add_read(read) == r.repeat_add_read(read)
        assert g.input_novel_read("IGHV1-2*01", read, -1, -1) == r.input_novel_read("IGHV1-2*01", read, -1, -1)
        assert g.add_read(read, "IGHV", 0, -1, 1, 0, 0.9) == r.add_read(read, "IGHV", 0, -1, 1, 0, 0.9)
        assert g.repeat_add_read(read) == r.repeat_add_read(read)
    assert g.output() == r.output()
    assert g.index_checksum() == r.index_checksum()
    # zero-length batch
    e = synth.READ_DESC
    assert g.run_descs(synth.run_cfg(), np.zeros(0, dtype=e), np.zeros(16, dtype=np.uint8), [])[0] == 0


def test_gpu_big_repeats(gpu_lib, ref):
    """k-mers with > 10000 postings (the `repeats` rules and their indexing slip, SeqSet.hpp:798-808, 871-887, 931-947)."""
    pc.check_big_repeats(gpu_lib, ref)


def test_gpu_barcode_mode(gpu_lib, ref):
    pc.check_barcode_mode(gpu_lib, ref)


def test_gpu_barcode_release(gpu_lib, ref):
    """ReleaseFinishedBarcodeSeq / ReleaseShallowContigs on the device (batch loop and per-call), contigMinCov 0/2/20."""
    pc.check_barcode_release(gpu_lib, ref)


def test_gpu_input_novel_fa(gpu_lib, ref, tmp_path):
    pc.check_input_novel_fa(gpu_lib, ref, tmp_path)


def test_gpu_probe_batch(gpu_lib, ref):
    """The dedicated warp-per-read probe kernel (t4_probe_kernel) vs the reference's GetHitsFromRead."""
    pc.check_probe_batch(gpu_lib, ref, seed=41, n_shards=5)
    pc.check_probe_batch(gpu_lib, ref, seed=42, n_shards=1, sample=120)


def test_gpu_single_cell_streams(gpu_lib, ref):
    """configs[3] in small: barcode-partitioned streams with per-barcode purge, on the device."""
    pc.check_single_cell(gpu_lib, ref, n_barcodes=60, reads_per_barcode=400, n_shards=7)
    pc.check_single_cell(gpu_lib, ref, seed=52, n_barcodes=9, reads_per_barcode=260, n_shards=2, contig_min_cov=3)


def test_gpu_repseq_streams(gpu_lib, ref):
    """configs[4] in small: repetitiveData = true (allowTotalSkip pass, mismatch factor 2.0), pseudo barcodes."""
    pc.check_repseq(gpu_lib, ref, n_reads=12000, n_shards=4)


def test_gpu_dup_runs(gpu_lib, ref):
    """Run-length RepeatAddRead in the device loop vs the reference, with consensus updates and k changes inside dup runs."""
    pc.check_dup_runs(gpu_lib, ref)


def test_gpu_batch_gene_grouped_shards(gpu_lib, ref):
    """Streams built by grouping reads by annotated gene (bench.py --shard-by gene): per-shard parity as for any other sharding."""
    assert pc.check_batch_vs_ref(gpu_lib, ref, 8, 9, nclones=40, npairs=700, group="gene") > 100


@pytest.mark.parametrize("seed,shards,kmer,drop", [(81, 1, 17, 0.0), (82, 4, 17, 0.15), (83, 40, 19, 0.0)])
def test_gpu_assign_pass(gpu_lib, ref, seed, shards, kmer, drop):
    """SURVEY.md 8f-2 on the device (t4_aux_kernel): extended sets by InputSeqSet at k, AssignRead of every assembled read
    by worker CTAs over the whole GPU, RecomputePosWeight -- assignments (contig, coordinates, strand, matchCnt, similarity
    double) and the extended sets' Output / index equal to the reference's pass (main.cpp:2047-2118)."""
    listed, assigned = pc.check_assign_pass(gpu_lib, ref, seed, shards, nclones=60, npairs=2500, kmer=kmer, drop=drop)
    assert listed > 2000 and assigned > 1500


def test_gpu_assign_pass_noisy_and_duplicates(gpu_lib, ref):
    cl = synth.make_clones(60, 92)
    w = synth.build_workload(cl, synth.sample_pairs(cl, 1200, 150, 92, sub_rate=0.04))
    listed, assigned = pc.check_assign_pass(gpu_lib, ref, 92, 2, workload=w)
    assert listed - assigned > 50
    cl = synth.make_clones(9, 95, chains=("TRB",))
    w = synth.build_workload(cl, synth.sample_amplicon(cl, 5000, 100, 95, alpha=0.7, sub_rate=0.002), repseq=True)
    pc.check_assign_pass(gpu_lib, ref, 95, 2, workload=w, cfg=synth.run_cfg(repetitive=1, first_read_len=100))
    # fewer workers than sets, and a single worker
    pc.check_assign_pass(gpu_lib, ref, 84, 6, nclones=30, npairs=500, n_workers=3)
    pc.check_assign_pass(gpu_lib, ref, 85, 2, nclones=30, npairs=500, n_workers=1)


@pytest.mark.parametrize("k", [21, 9, 31])
def test_gpu_kmer_count_stats(gpu_lib, ref, k):
    """SURVEY.md 8f-3 on the device (t4_kcount_kernel): canonical k-mer counts in one HBM hash table, per-read min / median /
    avg equal to the reference's KmerCount::AddCount + GetCountStatsAndTrim (ragged reads, N's, duplicates)."""
    assert pc.check_kmer_count_stats(gpu_lib, ref, seed=100 + k, n=6000, k=k) >= 6000


@pytest.mark.parametrize("seed,radius,hit_len,k", [(121, None, 27, 9), (122, 0, 23, 9), (124, 5, 30, 11)])
def test_gpu_refset_scan(gpu_lib, ref, tmp_path, seed, radius, hit_len, k):
    """SURVEY.md 8f-4 on the device: the reference gene set by InputRefFa (names, index) and fastq-extractor's per-read
    predicate -- IsLowComplexity and HasHitInSet(read, 0) with the reference-sequence chain rules (radius windows, LIS) --
    equal to the reference for 3000 reads (candidates of both strands, random reads, chimeras, indels, N's, short reads)."""
    assert pc.check_refset_scan(gpu_lib, ref, tmp_path, seed=seed, n=3000, radius=radius, hit_len=hit_len, k=k) > 500
