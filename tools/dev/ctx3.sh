for subs in config3 config5_weak,config3 config3,config4p; do
timeout 300 python bench.py --steps 20 --no-cpu --e2e-steps 1 --subs $subs 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SUB', {k:round(v['ms_per_step'],4) for k,v in d['sub_records'].items()})"
done
