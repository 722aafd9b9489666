# usage: bench_wl.sh workload [reps]
for i in $(seq 1 ${2:-3}); do timeout 200 python bench.py --workload $1 --steps 30 --no-cpu --no-sub --e2e-steps 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MS', d['config']['name'], d['ms_per_step'], d['roofline']['frac'])"; done
