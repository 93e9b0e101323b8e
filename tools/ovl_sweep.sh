run() { python bench.py --steps 20 --warmup 5 --no-post --no-strong --no-cpu $3 --chains $1 2>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', $1, d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('moving') or {}).get('value'))"; grep "e2e phases" gpurun_out/sweep.err; }
run 256 default
run 148 default --no-moving
run 32 default --no-moving
