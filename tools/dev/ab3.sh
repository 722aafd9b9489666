timeout 300 python -m pytest tests -x -q -m gpu -k "linearize or config3 or config5 or apply or golden or smoke" 2>&1 | tail -3
run() { timeout 200 python bench.py --steps 30 --no-cpu --no-sub --e2e-steps 1 --workload config3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MS', '$1', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; }
for i in 1 2; do SHDR_LIB=$PWD/tools/_build/libshdr_A.so run A; run B; done
