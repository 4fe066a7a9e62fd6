"""Golden vectors for the encoding fast path: the REAL reference DataHandler
(/root/reference/src/data_handler.py:27-71,101-141) on the column-dtype cases of
tests/util.dtype_frames.  Run in the build container only:

    python tests/golden/make_golden_encoding.py

Writes tests/golden/encoding_dtypes.npz (+ .json for the dictionaries)."""
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
    from data_handler import DataHandler
    from tests.util import dtype_frames
    logging.getLogger("MMSBM").setLevel(logging.ERROR)
    arrays, dicts = {}, {}
    for name, (train, test) in dtype_frames().items():
        dh = DataHandler()
        arrays[name + "_train"] = dh.format_train_data(train.copy())
        arrays[name + "_test"] = dh.format_test_data(test.copy())
        dicts[name] = [list(d.items()) for d in dh.return_dicts()]      # insertion order kept
    np.savez_compressed(os.path.join(HERE, "encoding_dtypes.npz"), **arrays)
    json.dump(dicts, open(os.path.join(HERE, "encoding_dtypes.json"), "w"))
    print({k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
