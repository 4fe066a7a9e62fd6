"""mmsbm_b200 -- B200-native EM hot path behind the API of eudald-seeslab/mmsbm.

    from mmsbm_b200 import MMSBM
    model = MMSBM(user_groups=20, item_groups=20, iterations=400, sampling=8, seed=1)
    model.fit(train_df); model.predict(test_df); model.score()

Compute lives in libmmsbm_b200.so (hand-written sm_100a CUDA, C ABI in
include/mmsbm_b200.h); build it with ``python -m mmsbm_b200.build``.
"""
from .backend import load_backend
from .data_handler import DataHandler
from .expectation_maximization import ExpectationMaximization
from .mmsbm import MMSBM

__all__ = ["MMSBM", "ExpectationMaximization", "DataHandler", "load_backend"]
__version__ = "0.1.0"
