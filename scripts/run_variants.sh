# usage: bash scripts/run_variants.sh "A:1 A:0 E:0 ..."   (variant letter : VLFB_FUSE_GRAD_FINISH)
for vf in $1; do v=${vf%%:*}; f=${vf##*:}
  VLFB_LIB=$PWD/gpurun_variants/libvlfb_$v.so VLFB_FUSE_GRAD_FINISH=$f python bench.py --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v fuse $f: %.3f ms/step  %.1f clips/s  e2e %.1f  sm %s MHz' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))"
done
