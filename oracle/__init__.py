"""CPU oracle for the LFB hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product path (the `vlfb` package and libvlfb.so)
never imports it and has no CPU fallback.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 and its arithmetic lives in Caffe2, which is absent from
/root/reference and from this image.  This package restates the reference graph
(lib/models/*.py) on PyTorch-CPU fp32/fp64 and the Caffe2 operator semantics from
their published definitions; every function cites the reference file:line it follows.
"""
