# Last GPU run of round 2 (one box, N = 1): the reworked k-mer counting kernels (tests), the default bench line, and the
# drop-in CLI comparison with the AssignRead pass on the device.
O=gpurun_out/r2l
mkdir -p $O
(timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kmer_count or assign" > $O/pytest_new.log 2>&1); echo "tests rc=$?"; tail -3 $O/pytest_new.log
(timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); echo "bench rc=$?"; tail -c 300 $O/bench_n1.err
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r2l/bench_n1.json'))
    print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value']), 'parity', (d.get('parity_spot_check') or {}).get('equal_reference'))
    a = d.get('assign_pass') or {}
    print('assign', a.get('ms'), a.get('reads_per_s'), (a.get('parity_spot_check') or {}).get('equal_reference'), a.get('error'))
    print('kmer_stats', json.dumps(d.get('preprocess_kmer_stats')))
except Exception as e:
    print('bench line ERR', e)
P
(timeout 200 python bench/cli_compare.py --pairs 30000 --streams 1,64 --shard-by gene > $O/cli_compare.json 2> $O/cli_compare.err); echo "cli rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2l/cli_compare.json'))
for r in d['runs']: print(r['binary'], round(r['wall_s'],1), r['addread_loop_s_from_log'], r.get('assign_pass_on_device'), r.get('identical_to_stock'))
"
