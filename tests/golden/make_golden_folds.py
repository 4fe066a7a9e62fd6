"""Golden fold indices: the REAL reference's cv_fit (/root/reference/src/mmsbm.py:371-472) with
fit / predict / score stubbed out, so only its fold construction runs (it is the part that
consumes ``self.rng``).  Records, per case of tests/util.fold_frames and per fold, the index
labels of the test frame in order, a SHA-1 of the train frame's labels (joined by commas, in
order), and the next draw of the generator afterwards.
Run in the build container only:

    python tests/golden/make_golden_folds.py   ->  tests/golden/folds.json"""
import hashlib
import json
import logging
import os
import sys

import numpy as np

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def main():
    assert os.path.isdir(REF), "reference not mounted; run this in the build container"
    sys.path.insert(0, REF)
    os.chdir("/tmp")
    from mmsbm import MMSBM
    from tests.util import fold_frames
    logging.getLogger("MMSBM").setLevel(logging.ERROR)
    out = {}
    for name, (frame, folds) in fold_frames().items():
        model = MMSBM(2, 2, iterations=1, seed=7, backend="numpy")
        seen = {"train": [], "test": []}
        model.fit = lambda train, silent=False: seen["train"].append([str(a) for a in train.index])
        model.predict = lambda test: seen["test"].append([str(a) for a in test.index]) or np.zeros((1, 1))
        model.score = lambda silent=False: {"stats": {"accuracy": 0.0}, "objects": {}}
        model.cv_fit(frame, folds=folds)
        out[name] = {"train_sha1": [hashlib.sha1(",".join(t).encode()).hexdigest() for t in seen["train"]],
                     "test": seen["test"],
                     "next_random": float(model.rng.random())}
    json.dump(out, open(os.path.join(HERE, "folds.json"), "w"))
    print({k: [len(t) for t in v["test"]] for k, v in out.items()})


if __name__ == "__main__":
    main()
