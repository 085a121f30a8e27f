"""Seeded synthetic inputs — re-exported from the package so tests, bench.py
and smoke() share one generator."""
from refinedet.pytorch_b200.synthetic import *  # noqa: F401,F403
from refinedet.pytorch_b200.synthetic import (assert_tie_free, detect_inputs, detect_inputs_clustered,  # noqa: F401
                                              detect_logits, targets,
                                              tie_free_detect_inputs, train_predictions)
