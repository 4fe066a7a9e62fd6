"""Fold helpers with the behaviour of the reference's src/helpers.py:4-24."""


def _invert_dict(d):
    return {v: k for k, v in d.items()}


def get_n_per_group(x, n, rng):
    """Up to ``n`` index labels of group ``x`` drawn without replacement; when the
    group is smaller than ``n`` the request shrinks until it fits (helpers.py:8-13).
    ``Generator.choice`` validates the size before it draws, so a failed attempt
    consumes no random numbers and asking directly for min(n, len) is equivalent."""
    take = min(n, len(x.index))
    if take <= 0:
        return None
    return rng.choice(x.index, take, replace=False)


def structure_folds(data, folds):
    """Rows held out per user and fold = int(n_items / folds) (helpers.py:16-24)."""
    n_items = len(set(data.iloc[:, 1]))
    assert folds <= n_items, (
        f"Fold number can't be higher than {n_items} since this is the number of "
        f"different items you have.")
    return int(n_items / folds)
