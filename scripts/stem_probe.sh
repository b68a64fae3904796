for cfg in "2048 128" "128 2048"; do set -- $cfg
  echo "== LBO=$1 SBO=$2"; VLFB_STEM_IM2COL=1 VLFB_STEM_LBO=$1 VLFB_STEM_SBO=$2 timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "stem_conv and tcgen05" 2>&1 | grep -E "passed|failed|assert|Error" | head -4
done
for sw in 1 0 1; do VLFB_STEM_IM2COL=$sw python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/stem_err$sw.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stem_im2col $sw: %.3f ms/step  %.1f clips/s loss %s' % (d['ms_per_step'], d['value'], d['loss']))"; done
for sw in 1 0; do VLFB_SIDE_WGRAD=$sw python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/side_err$sw.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('side_wgrad $sw: %.3f ms/step  %.1f clips/s loss %s' % (d['ms_per_step'], d['value'], d['loss']))"; done
