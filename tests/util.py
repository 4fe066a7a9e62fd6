"""Seeded synthetic problems shared by the tests and the bench (no reference needed)."""
import numpy as np
import pandas as pd


def mock_data(seed, n=100):
    """The generator of the reference's own fixture (tests/test_mmsbm.py:12-22)."""
    rng = np.random.default_rng(seed)
    return pd.DataFrame({
        "users": [f"user{rng.choice(list(range(5)))}" for _ in range(n)],
        "items": [f"item{rng.choice(list(range(10)))}" for _ in range(n)],
        "ratings": [rng.choice(list(range(1, 6))) for _ in range(n)],
    })


def random_triples(seed, N, U, I, R, heavy_tail=False):
    """int64 [N,3]; every user / item / rating id appears at least once (N >= max(U,I,R))."""
    g = np.random.default_rng(seed)
    if heavy_tail:   # Zipf-like popularity, exponent ~1
        pu = 1.0 / np.arange(1, U + 1); pu /= pu.sum()
        pi = 1.0 / np.arange(1, I + 1); pi /= pi.sum()
        u = g.choice(U, size=N, p=pu)
        i = g.choice(I, size=N, p=pi)
    else:
        u = g.integers(0, U, N)
        i = g.integers(0, I, N)
    r = g.integers(0, R, N)
    u[:U] = np.arange(U); i[:I] = np.arange(I); r[:R] = np.arange(R)
    return np.stack([u, i, r], axis=1).astype(np.int64)


def random_params(seed, U, I, K, L, R, S=None):
    g = np.random.default_rng(seed)
    lead = () if S is None else (S,)
    theta = g.random(lead + (U, K)); eta = g.random(lead + (I, L)); pr = g.random(lead + (K, L, R))
    theta /= theta.sum(axis=-1, keepdims=True)
    eta /= eta.sum(axis=-1, keepdims=True)
    pr /= pr.sum(axis=-1, keepdims=True)
    return theta, eta, pr


def rel_err(a, b):
    """max over elements of |a-b| / max(|b|, tiny): the per-element relative error the
    north_star tolerance (1e-10) is stated in."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def dtype_frames(seed=1, n=3000):
    """(train, test) frames per column-dtype case for the encoding tests: integer, float (with
    -0.0 / inf / float32), str, bool / small ints, mixed python objects / nullable / categorical,
    datetimes, NaN.  Each test frame holds a sample of the train rows plus rows with unseen ids."""
    g = np.random.default_rng(seed)
    cases = {
        "int": pd.DataFrame({"users": g.integers(-50, 300, n), "items": g.integers(0, 120, n),
                             "ratings": g.integers(1, 6, n)}),
        "float": pd.DataFrame({"users": g.integers(0, 300, n).astype(float),
                               "items": g.choice([0.0, -0.0, 1.5, 2.25, 1e20, np.inf], n),
                               "ratings": g.choice([0.5, 1.0, 1.5], n).astype(np.float32)}),
        "str": pd.DataFrame({"users": ["u%d" % x for x in g.integers(0, 300, n)],
                             "items": ["%d" % x for x in g.integers(0, 120, n)],
                             "ratings": [str(x) for x in g.integers(1, 6, n)]}),
        "bool": pd.DataFrame({"users": g.integers(0, 2, n).astype(bool), "items": g.integers(0, 120, n).astype(np.uint8),
                              "ratings": g.integers(1, 6, n).astype(np.int16)}),
        "mixed": pd.DataFrame({"users": pd.Series(g.choice(np.array([1, 1.0, True, "1", "a", 2, 2.5], dtype=object), n),
                                                  dtype=object),
                               "items": pd.Series(g.integers(0, 50, n)).astype("Int64"),
                               "ratings": pd.Categorical(g.choice(["lo", "mid", "hi"], n))}),
        "dates": pd.DataFrame({"users": pd.to_datetime(g.integers(0, 40, n), unit="D"),
                               "items": g.integers(0, 30, n), "ratings": g.integers(1, 4, n)}),
        "nan": pd.DataFrame({"users": g.choice([1.0, np.nan, 3.0], n), "items": g.integers(0, 30, n),
                             "ratings": g.integers(1, 4, n)}),
    }
    out = {}
    for name, d in cases.items():
        extra = d.iloc[:3].copy()
        if name == "str":
            extra.iloc[0, 1] = "100000"                       # an item that is not in the train set
        elif name == "bool":
            extra.iloc[0, 1] = 200                            # items are 0..119 (uint8)
        elif name != "mixed":
            extra.iloc[0, 1] = extra.iloc[0, 1] + 100000
        out[name] = (d, pd.concat([d.iloc[::7], extra]))
    return out


def fold_frames(seed=0):
    """(frame, folds) cases for the fold-construction tests: the reference fixture, integer / str /
    float users, a shuffled index and an index of strings that contains the label "0"."""
    g = np.random.default_rng(seed)
    n = 6000
    return {
        "fixture": (mock_data(1), 2),
        "ints": (pd.DataFrame({"users": g.integers(0, 300, n), "items": g.integers(0, 40, n),
                               "ratings": g.integers(1, 6, n)}), 5),
        "strs": (pd.DataFrame({"users": ["u%d" % x for x in g.integers(0, 300, n)], "items": g.integers(0, 23, n),
                               "ratings": g.integers(1, 6, n)}), 4),
        "shuffled_index": (pd.DataFrame({"users": g.integers(0, 50, 2000), "items": g.integers(0, 30, 2000),
                                         "ratings": g.integers(1, 6, 2000)}).sample(frac=1.0, random_state=3), 3),
        "str_index": (pd.DataFrame({"users": g.integers(0, 50, 500), "items": g.integers(0, 12, 500),
                                    "ratings": g.integers(1, 6, 500)}, index=[str(i) for i in range(500)]), 3),
        "floats": (pd.DataFrame({"users": g.integers(0, 50, 2000).astype(float), "items": g.integers(0, 30, 2000),
                                 "ratings": g.integers(1, 6, 2000)}), 3),
    }
