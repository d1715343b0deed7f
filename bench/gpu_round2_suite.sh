# Round-2 measurement suite (one B200): parity tests, the default bench line (gene streams), launch-order and configs[3]/[4]
# runs, the reference arm, the drop-in CLI comparison, the ncu launch list and one full capture each of the stream and the
# probe kernel.  Outputs under gpurun_out/r2final/; copy what matters to profiles/.
O=gpurun_out/r2final
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > $O/pytest.log; cat $O/pytest.log
(timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); tail -c 200 $O/bench_n1.err
(T4_STREAM_ORDER=index timeout 420 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality --no-probe > $O/bench_index_order.json 2> $O/bench_index_order.err); tail -c 200 $O/bench_index_order.err
(timeout 600 python bench.py --config 3 --steps 2 --warmup 1 > $O/bench_cfg3.json 2> $O/bench_cfg3.err); tail -c 300 $O/bench_cfg3.err
(timeout 600 python bench.py --config 4 --steps 2 --warmup 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err); tail -c 300 $O/bench_cfg4.err
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err)
(timeout 900 python bench/cli_compare.py --pairs 30000 --streams 1,64,1024 > $O/cli_compare.json 2> $O/cli_compare.err); tail -c 300 $O/cli_compare.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality > $O/launches.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:t4_stream_kernel -s 1 -c 1 -o $O/stream python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-quality --no-probe > $O/ncu_stream.log 2>&1); tail -2 $O/ncu_stream.log
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:t4_probe_kernel -s 1 -c 1 -o $O/probe python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-quality > $O/ncu_probe.log 2>&1); tail -2 $O/ncu_probe.log
python -c "
import json
for f in ('n1','index_order','cfg3','cfg4','reference_arm'):
    try:
        d=json.load(open('$O/bench_%s.json'%f)); p=d.get('roofline_probe'); q=d.get('assembly_quality'); print(f, round(d['value']), round(d['e2e']['value']), d.get('cpu_baseline') and round(d['cpu_baseline']['value']), d.get('parity_spot_check'), p and (p['kernel_ms'], round(p['frac'],3), round(p['frac_with_16B_hits'],3)), q and round(q['spanned_fraction'],4), d.get('roofline') and d['roofline'].get('kernel_ms'))
    except Exception as e: print(f, 'ERR', e)
try:
    d=json.load(open('$O/cli_compare.json'))
    for r in d['runs']: print(r['binary'], round(r['wall_s'],1), r.get('addread_loop_s_from_log'), r.get('identical_to_stock'), r.get('contiguity'))
except Exception as e: print('cli ERR', e)
"
