O=gpurun_out/r2o
mkdir -p $O
(timeout 600 python -m pytest tests -m gpu -q -k "probe or gene or hits" 2>&1 | tail -5) > $O/pytest_probe.log; cat $O/pytest_probe.log
(timeout 420 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality > $O/bench_redux.json 2> $O/bench_redux.err); tail -c 300 $O/bench_redux.err
python -c "
import json
d=json.load(open('$O/bench_redux.json')); p=d['roofline_probe']; print(round(d['value']), round(d['e2e']['value']), p['kernel_ms_all'], round(p['frac'],4), round(p['frac_with_16B_hits'],4))
"
