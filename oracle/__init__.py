"""
ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT.

A CPU (NumPy float32) restatement of the per-anchor detection hot path of
TropComplique/single-shot-detector, written op for op (same temporaries, same
evaluation order, float32 everywhere) from the reference sources cited in each
function.  It exists so that the CUDA path can be checked against the
reference's semantics on identical inputs.

Who may import this package: tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package
(single-shot-detector_b200/) never imports it and has no CPU fallback.

How it is pinned.  The reference has no tests, golden vectors or fixtures for
this path and its only runtime, TensorFlow 1.12, cannot be
installed here.  Instead, tests/golden/make_golden.py executes the reference's
OWN unmodified Python files from /root/reference on top of
oracle/tf_numpy_shim (an eager NumPy stand-in for the ~60 TF symbols the path
uses) and freezes inputs + outputs as fixtures under tests/golden/;
tests/test_oracle_golden.py checks this restatement against them bit for bit
(transcendental-free outputs) or to 1 ulp-level tolerances.  Everything written in
the reference tree is therefore pinned.  What remains unpinned is TensorFlow's
own kernel arithmetic, which is not in the reference tree:

  * tf.image.non_max_suppression (NonMaxSuppressionV3, TF 1.12, called at
    detector/utils/nms.py:33) -- restated from the published algorithm in
    oracle/nms.py::non_max_suppression_v3 and in oracle/csrc/oracle_nms.c;
    cross-checked against torchvision.ops.nms (CPU) in tests.  PARITY UNPINNED
    for this op (no reference-side golden vector exists).
  * last-ulp behaviour of Eigen's exp/log/log1p/sigmoid/pow.

Conventions: boxes are [ymin, xmin, ymax, xmax], float32, normalised; indices
int32; `matches` uses -1 = background, -2 = ignore.
"""
