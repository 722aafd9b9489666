# A/B two builds of libshdr in ONE gpurun call (boxes differ by ~1 %): A = in-tree, B = tools/_build/libshdr_B.so
run() { timeout 200 python bench.py --steps 30 --no-cpu --no-sub --e2e-steps 1 --workload ${WL:-config2} 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MS', '$1', d['ms_per_step'], d['roofline']['frac'])"; }
for i in 1 2; do run A; SHDR_LIB=$PWD/tools/_build/libshdr_B.so run B; done
