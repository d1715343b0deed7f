# Round-2 ncu evidence for the kernels added in this session (one box, N = 1): launch list of a default bench run and one
# `--set full` capture each of the three t4_aux_kernel launches (PREP, ASSIGN, RECOMPUTE) and the two t4_kcount_kernel launches.
O=gpurun_out/r2n
mkdir -p $O
(timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kmer_count" > $O/pytest_kmer.log 2>&1); echo "kmer tests rc=$?"; tail -3 $O/pytest_kmer.log
(timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-quality > $O/launches.log 2>&1); echo "launch list rc=$?"; tail -c 300 $O/launches.log
(timeout 480 ncu --set full --clock-control none --import-source on -k "regex:t4_aux_kernel|t4_kcount_kernel" -c 8 -o $O/aux python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-quality --no-probe > $O/ncu_aux.log 2>&1); echo "full capture rc=$?"; tail -3 $O/ncu_aux.log
ls -la $O
