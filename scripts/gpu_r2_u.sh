#!/bin/bash
# call U: last sanity check of the final tree (smoke + tiny training-step parity)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2u_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/r2u_smoke.log
timeout 60 python -m pytest tests/test_gpu_model.py -q -k "tiny_fbo_nl_train_step or grad_finish_fusion" > $O/r2u_tiny.log 2>&1; echo "tiny rc=$?"; tail -n 2 $O/r2u_tiny.log
