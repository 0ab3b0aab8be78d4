"""Constants of the hot path; same names and values as reference detector/constants.py."""

# for fpn only: the minimal feature stride is 2**MIN_LEVEL (constants.py:4)
MIN_LEVEL = 3
DIVISOR = 128                                   # constants.py:7

# layout of the per-level tower outputs consumed by box_predictor.reshape_and_concatenate (constants.py:9)
DATA_FORMAT = 'channels_first'

EPSILON = 1e-8                                  # constants.py:12 (compiled into the kernels as SSDK_EPS)
SCALE_FACTORS = [10.0, 10.0, 5.0, 5.0]          # constants.py:15 (compiled into box_encode / box_decode)

# thresholds for IoU when creating training targets (constants.py:25-26)
POSITIVES_THRESHOLD = 0.5
NEGATIVES_THRESHOLD = 0.5

# the reference runs at most 8 images concurrently in tf.map_fn (constants.py:29); here all images of a
# batch are processed by the same kernel launch, the constant is kept for API completeness only
PARALLEL_ITERATIONS = 8
