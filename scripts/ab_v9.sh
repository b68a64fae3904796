run() { ( cd $1; VLFB_FUSE_GRAD_FINISH=$2 python bench.py --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3: %.3f ms/step  %.1f clips/s  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))" ); }
run gpurun_variants/v9 0 v9
run . 0 cur_fuse0
run . 1 cur_fuse1
run gpurun_variants/v9 0 v9
run . 1 cur_fuse1
