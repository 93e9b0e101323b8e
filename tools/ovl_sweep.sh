run() { python bench.py --steps 20 --warmup 5 --no-post --no-moving --no-strong --no-cpu --chains $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', $1, d['value'], d['ms_per_step'])"; }
run 256 default
RCB200_TW_SMEM=1 run 256 twsmem
run 148 default
RCB200_TW_SMEM=0 run 148 twglobal
