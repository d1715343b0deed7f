O=gpurun_out/r2n
mkdir -p $O
(timeout 420 python bench.py --shard-by gene --streams 2048 --steps 2 --warmup 1 --no-cpu-baseline --dump-streams $O/streams_gene2048.npz > $O/bench_gene2048.json 2> $O/bench_gene2048.err); tail -c 300 $O/bench_gene2048.err
(timeout 420 python bench.py --shard-by gene --streams 4096 --steps 2 --warmup 1 --no-cpu-baseline --dump-streams $O/streams_gene4096.npz > $O/bench_gene4096.json 2> $O/bench_gene4096.err); tail -c 300 $O/bench_gene4096.err
python -c "
import json
for f in ('gene2048','gene4096'):
    try:
        d=json.load(open('$O/bench_%s.json'%f)); print(f, round(d['value']), round(d['e2e']['value']), d.get('parity_spot_check'), d.get('assembly_quality'), d.get('stream_balance'))
    except Exception as e: print(f, 'ERR', e)
"
