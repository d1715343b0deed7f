# The very last GPU call of round 2: the whole GPU suite (38 tests, incl. the stage-0 scan) and the stage-0 scan figure.
O=gpurun_out/r2e
mkdir -p $O
(timeout 110 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1); echo "gpu suite rc=$?"; tail -4 $O/pytest_gpu.log
(timeout 60 python bench/stage0_scan.py --reads 1000000 --sample 8000 > $O/stage0_scan.json 2> $O/stage0_scan.err); echo "stage0 rc=$?"; cat $O/stage0_scan.json | head -c 1500; tail -c 300 $O/stage0_scan.err
