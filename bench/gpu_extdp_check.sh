O=gpurun_out/r2p
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > $O/pytest.log; cat $O/pytest.log
(timeout 420 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality --no-probe > $O/bench_warpdp.json 2> $O/bench_warpdp.err); tail -c 300 $O/bench_warpdp.err
(T4_LIB_PATH=$PWD/trust4_b200/libtrust4_b200_threaddp.so timeout 420 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality --no-probe > $O/bench_threaddp.json 2> $O/bench_threaddp.err); tail -c 300 $O/bench_threaddp.err
python -c "
import json
for f in ('warpdp','threaddp'):
    d=json.load(open('$O/bench_%s.json'%f)); r=d['roofline']; print(f, round(d['value']), round(d['e2e']['value']), r['kernel_ms'], r['phase_share'], r.get('extend_split'), r['stream_balance']['mean_ms'])
"
