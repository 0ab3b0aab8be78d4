"""Oracle constants. Follows reference detector/constants.py:12,15,25-26,29."""
import numpy as np

EPSILON = np.float32(1e-8)                      # constants.py:12
SCALE_FACTORS = [10.0, 10.0, 5.0, 5.0]          # constants.py:15
POSITIVES_THRESHOLD = 0.5                       # constants.py:25
NEGATIVES_THRESHOLD = 0.5                       # constants.py:26
PARALLEL_ITERATIONS = 8                         # constants.py:29
