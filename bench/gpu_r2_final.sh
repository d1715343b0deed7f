# Round-2 closing GPU run (one box, N = 1): the new AssignRead pass first (tests, then the bench line that carries it),
# then the whole GPU suite.  Everything lands in gpurun_out/r2f/.
O=gpurun_out/r2f
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/gpu.txt 2>&1
(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "assign or kmer_count" > $O/pytest_assign.log 2>&1); echo "assign tests rc=$?"; tail -4 $O/pytest_assign.log
(timeout 480 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); echo "bench rc=$?"; tail -c 400 $O/bench_n1.err
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r2f/bench_n1.json'))
    p = d.get('roofline_probe') or {}
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value']),
          'parity', d.get('parity_spot_check'), 'probe', p.get('kernel_ms'), p.get('frac'), 'stream_ms', d['roofline'].get('kernel_ms'), 'frac', d['roofline'].get('frac'))
    print('assign_pass', json.dumps(d.get('assign_pass')))
    print('kmer_stats', json.dumps(d.get('preprocess_kmer_stats')))
    print('quality', d.get('assembly_quality'))
except Exception as e:
    print('bench line ERR', e)
P
(timeout 700 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1); echo "gpu suite rc=$?"; tail -6 $O/pytest_gpu.log
