"""slslam_b200 — B200-native solver for SLSLAM's line bundle adjustment and pose-graph optimisation.

The product is the C-ABI shared library built from `csrc/` (see `include/slslam_b200.h`); this Python
package is only the loader/binding used by the tests and the benchmark, plus the synthetic-input generator.
"""
__version__ = "0.1.0"
