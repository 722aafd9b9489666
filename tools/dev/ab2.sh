# tests with the in-tree build, then A/B: A = tools/_build/libshdr_A.so (previous commit), B = in-tree
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "pooled or frontend or nan or constant or sparse or config2" 2>&1 | tail -3
run() { timeout 200 python bench.py --steps 30 --no-cpu --no-sub --e2e-steps 1 --workload ${WL:-config2} 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MS', '$1', d['ms_per_step'], d['roofline']['frac'])"; }
for i in 1 2; do SHDR_LIB=$PWD/tools/_build/libshdr_A.so run A; run B; done
