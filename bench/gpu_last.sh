O=gpurun_out/r2last
mkdir -p $O
(timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); tail -c 200 $O/bench_n1.err
(T4_BENCH_CFG4_SHARD_BY=gene timeout 600 python bench.py --config 4 --steps 2 --warmup 1 > $O/bench_cfg4_gene.json 2> $O/bench_cfg4_gene.err); tail -c 300 $O/bench_cfg4_gene.err
(timeout 600 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_cfg4_rank.json 2> $O/bench_cfg4_rank.err); tail -c 300 $O/bench_cfg4_rank.err
python -c "
import json
for f in ('n1','cfg4_gene','cfg4_rank'):
    try:
        d=json.load(open('$O/bench_%s.json'%f)); p=d.get('roofline_probe'); q=d.get('assembly_quality'); print(f, round(d['value']), round(d['e2e']['value']), d.get('cpu_baseline') and round(d['cpu_baseline']['value']), d.get('parity_spot_check'), p and (p['kernel_ms'], round(p['frac'],3), round(p['frac_with_16B_hits'],3)), q, d['contigs_per_gpu'], d.get('roofline') and d['roofline'].get('kernel_ms'))
    except Exception as e: print(f, 'ERR', e)
"
