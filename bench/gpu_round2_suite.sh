mkdir -p gpurun_out/r2g
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2g/pytest.log; cat gpurun_out/r2g/pytest.log
(timeout 900 python bench.py > gpurun_out/r2g/bench_n1.json 2> gpurun_out/r2g/bench_n1.err); tail -c 200 gpurun_out/r2g/bench_n1.err
for mb in 5 6 8; do (T4_LIB_PATH=$PWD/trust4_b200/libtrust4_b200_mb$mb.so timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality --no-probe > gpurun_out/r2g/bench_mb$mb.json 2> gpurun_out/r2g/bench_mb$mb.err); python -c "
import json
d=json.load(open('gpurun_out/r2g/bench_mb$mb.json')); print('mb$mb', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms'], d['roofline']['stream_balance']['max_ms'], d['roofline']['stream_balance']['mean_ms'])"; done
(timeout 600 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r2g/bench_cfg3.json 2> gpurun_out/r2g/bench_cfg3.err); tail -c 300 gpurun_out/r2g/bench_cfg3.err
(timeout 600 python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/r2g/bench_cfg4.json 2> gpurun_out/r2g/bench_cfg4.err); tail -c 300 gpurun_out/r2g/bench_cfg4.err
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality > gpurun_out/r2g/launches.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:t4_stream_kernel -s 1 -c 1 -o gpurun_out/r2g/stream python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-quality --no-probe > gpurun_out/r2g/ncu_stream.log 2>&1); tail -2 gpurun_out/r2g/ncu_stream.log
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:t4_probe_kernel -s 1 -c 1 -o gpurun_out/r2g/probe python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-quality > gpurun_out/r2g/ncu_probe.log 2>&1); tail -2 gpurun_out/r2g/ncu_probe.log
python -c "
import json
for f in ('n1','cfg3','cfg4'):
    try:
        d=json.load(open('gpurun_out/r2g/bench_%s.json'%f)); p=d['roofline_probe']; print(f, round(d['value']), round(d['e2e']['value']), d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['parity_spot_check'], p and (p['kernel_ms'], round(p['frac'],3), round(p['frac_with_16B_hits'],3)), d['roofline']['stream_balance']['max_ms'], d['roofline']['stream_balance']['mean_ms'])
    except Exception as e: print(f, 'ERR', e)
"
