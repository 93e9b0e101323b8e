run() { python bench.py --steps 20 --warmup 5 --no-post --no-moving --no-strong --no-cpu --chains $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', $1, d['value'], d['ms_per_step'])"; }
RCB200_RS_TEAM=160 run 32 team160
RCB200_RS_TEAM=192 run 32 team192
RCB200_RS_TEAM=256 run 32 team256
RCB200_RS_TEAM=160 run 148 team160
RCB200_RS_TEAM=192 run 148 team192
RCB200_OVERLAP_MIN_THREADS=256 RCB200_RS_TEAM=96 run 256 ovl96
RCB200_OVERLAP_MIN_THREADS=256 RCB200_RS_TEAM=64 run 296 ovl64
RCB200_OVERLAP_MIN_THREADS=256 RCB200_RS_TEAM=64 run 512 ovl64
run 512 base
