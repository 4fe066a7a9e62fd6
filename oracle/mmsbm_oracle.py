"""CPU oracle for the mmsbm EM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module restates, in plain numpy, the algorithm of the reference
(eudald-seeslab/mmsbm v1.0.7) for the path named by BASELINE.json:north_star.
It exists so that the CUDA path can be checked against something that runs on
the CPU.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may import it; nothing under
``mmsbm_b200/`` does, and the product path raises when the CUDA library is
missing instead of falling back to this file.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from /root/reference/src (in the build container, where it exists),
runs it on seeded inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those
vectors, and against the known answers of the reference's own test-suite
(tests/test_mmsbm.py:53-102, tests/test_backends.py:7-60).

Every function cites the reference ``file:line`` it follows (paths relative
to /root/reference).  Arithmetic is float64 throughout, ids are int64, exactly
as in the reference.
"""
from __future__ import annotations

import numpy as np

EPS = float(np.finfo(np.float64).eps)  # 2.220446049250313e-16, src/kernels_numpy.py:51


# --------------------------------------------------------------------------
# a1  omega[n,k,l] = theta[u_n,k] * eta[i_n,l] * pr[k,l,r_n]
# --------------------------------------------------------------------------
def omegas(data, theta, eta, pr):
    """Unnormalised responsibilities, shape [N,K,L].

    Follows src/kernels_numpy.py:21-36 (gather theta rows, gather eta rows,
    gather the [K,L] slab of the rating, multiply left to right)."""
    u, i, r = data[:, 0], data[:, 1], data[:, 2]
    slabs = np.moveaxis(pr, 2, 0)            # [R,K,L] view, as pr.transpose(2,0,1)
    out = theta[u][:, :, None] * eta[i][:, None, :]
    out = out * slabs[r]
    return out


# --------------------------------------------------------------------------
# a2  unnormalised M-step sums
# --------------------------------------------------------------------------
def em_sums(data, theta, eta, pr, chunk=None):
    """(n_theta[U,K], n_eta[I,L], n_pr[K,L,R]) -- src/kernels_numpy.py:43-79.

    inc = omega / max(sum_kl omega, eps); rows are scattered in data order
    (np.add.at is sequential), the pr slab of rating r is the sum over the
    rows with that rating and stays zero when no row has it (:74-77).

    ``chunk`` (rows per block) bounds the [N,K,L] temporaries for large N; it
    changes only the association of the pr sums (last bits).  ``None``
    reproduces the reference's operation order exactly.
    """
    n_theta = np.zeros_like(theta)
    n_eta = np.zeros_like(eta)
    n_pr = np.zeros_like(pr)
    n_rows = data.shape[0]
    step = n_rows if not chunk else int(chunk)
    for lo in range(0, max(n_rows, 1), max(step, 1)):
        block = data[lo:lo + step]
        if block.shape[0] == 0:
            break
        w = omegas(block, theta, eta, pr)
        tot = w.sum(axis=(1, 2))
        inc = w / np.maximum(tot, EPS)[:, None, None]
        np.add.at(n_theta, block[:, 0], inc.sum(axis=2))
        np.add.at(n_eta, block[:, 1], inc.sum(axis=1))
        for lvl in range(pr.shape[2]):
            sel = block[:, 2] == lvl
            if sel.any():
                if chunk:
                    n_pr[:, :, lvl] += inc[sel].sum(axis=0)
                else:
                    n_pr[:, :, lvl] = inc[sel].sum(axis=0)
    return n_theta, n_eta, n_pr


# --------------------------------------------------------------------------
# a3 / a4  normalisations
# --------------------------------------------------------------------------
def degree_factors(data, n_user_groups, n_item_groups):
    """Normalisation factors of src/mmsbm.py:100-111: max(degree,1) repeated
    over the group axis, int64 [U,K] and [I,L].  Ids are contiguous 0..max
    after encoding, so the dict-of-neighbour-lists of the reference reduces to
    a bincount."""
    n_users = int(data[:, 0].max()) + 1
    n_items = int(data[:, 1].max()) + 1
    du = np.maximum(np.bincount(data[:, 0], minlength=n_users), 1).astype(np.int64)
    di = np.maximum(np.bincount(data[:, 1], minlength=n_items), 1).astype(np.int64)
    return (np.repeat(du[:, None], n_user_groups, axis=1),
            np.repeat(di[:, None], n_item_groups, axis=1))


def scale_by_degree(x, factors):
    """src/expectation_maximization.py:118-120."""
    return x / factors


def normalize_pr(x):
    """Each pr[k,l,:] divided by its sum; a sum that is exactly zero is
    replaced by one (src/expectation_maximization.py:122-155)."""
    flat = x.reshape(-1, x.shape[2])
    tot = flat.sum(axis=1)
    tot = np.where(tot == 0, 1, tot)
    return (flat / tot[:, None]).reshape(x.shape)


# --------------------------------------------------------------------------
# a5  the quantity the reference calls "likelihood"
# --------------------------------------------------------------------------
def likelihood(data, theta, eta, pr):
    """sum_n sum_kl w~ (log w~ - log S~_n) with w~ = max(omega, eps) and
    S~ = max(sum_kl omega, eps) -- src/expectation_maximization.py:157-167."""
    w = omegas(data, theta, eta, pr)
    tot = np.zeros(data.shape[0])
    np.sum(w, axis=(1, 2), out=tot)
    ws = np.maximum(w, EPS)
    ts = np.maximum(tot, EPS)
    return np.sum(ws * np.log(ws) - ws * np.log(ts[:, None, None]))


# --------------------------------------------------------------------------
# a6  rating distribution of a (user,item) pair
# --------------------------------------------------------------------------
def rating_distribution(data, theta, eta, pr):
    """rat[n,r] = sum_kl theta[u_n,k] eta[i_n,l] pr[k,l,r]
    (src/kernels_numpy.py:86-97; the rating column of ``data`` is unused)."""
    return np.einsum('nk,nl,klr->nr', theta[data[:, 0]], eta[data[:, 1]], pr)


# --------------------------------------------------------------------------
# a7  one seeded EM run
# --------------------------------------------------------------------------
def child_seeds(seed, sampling):
    """src/mmsbm.py:82-85: children of the SeedSequence behind default_rng(seed)."""
    rng = np.random.default_rng(seed)
    return rng, rng.bit_generator._seed_seq.spawn(sampling)


def seeded_init(seed, n_users, n_items, K, L, R, fu, fi):
    """theta0, eta0, pr0 of src/mmsbm.py:224-233: three draws, in this order,
    from default_rng(seed); theta and eta divided by the DEGREE factors (not
    by their row sums), pr normalised over the rating axis."""
    rng = np.random.default_rng(seed)
    theta = scale_by_degree(rng.random((n_users, K)), fu)
    eta = scale_by_degree(rng.random((n_items, L)), fi)
    pr = normalize_pr(rng.random((K, L, R)))
    return theta, eta, pr


def em_iteration(data, theta, eta, pr, fu, fi, chunk=None):
    """One pass of the loop body, src/mmsbm.py:244-250."""
    nt, ne, npr = em_sums(data, theta, eta, pr, chunk=chunk)
    return scale_by_degree(nt, fu), scale_by_degree(ne, fi), normalize_pr(npr)


def run_em(data, seed, K, L, iterations, chunk=None):
    """src/mmsbm.py:187-269 without tqdm / logging.  Returns the same dict."""
    R = len(set(data[:, 2].tolist()))
    fu, fi = degree_factors(data, K, L)
    theta, eta, pr = seeded_init(seed, fu.shape[0], fi.shape[0], K, L, R, fu, fi)
    for _ in range(iterations):
        theta, eta, pr = em_iteration(data, theta, eta, pr, fu, fi, chunk=chunk)
    return {"likelihood": likelihood(data, theta, eta, pr),
            "pr": pr, "theta": theta, "eta": eta}


# --------------------------------------------------------------------------
# a8  the index structure (CSR by user / CSC by item)
# --------------------------------------------------------------------------
def bucket_order(primary, rating, n_primary, n_ratings):
    """Rows grouped by (primary id, rating), original order kept inside a
    bucket.  Returns (seg_ptr[n_primary*n_ratings+1], perm[N]) as int32.

    The reference holds the same information as per-id row lists
    ``np.where(train[:,c]==a)[0]`` (src/mmsbm.py:114-122).  The bucket of
    (a, r) is the ascending intersection of ``_user_indices[a]`` (or
    ``_item_indices[a]``) with ``_rating_indices[r]``; concatenating a's R
    buckets and sorting gives back the reference list -- checked in
    tests/test_oracle_golden.py."""
    key = primary.astype(np.int64) * n_ratings + rating.astype(np.int64)
    perm = np.argsort(key, kind="stable").astype(np.int32)
    counts = np.bincount(key, minlength=n_primary * n_ratings)
    seg_ptr = np.zeros(n_primary * n_ratings + 1, dtype=np.int64)
    np.cumsum(counts, out=seg_ptr[1:])
    return seg_ptr.astype(np.int32), perm


# --------------------------------------------------------------------------
# a10  prediction statistics and best-run choice
# --------------------------------------------------------------------------
def prediction_stats(rat, real, levels):
    """src/mmsbm.py:480-539.  ``levels`` are the ENCODED rating ids
    (sorted(set(train[:,2])), :95), ``real`` the encoded test ratings.
    Rows whose distribution sums to exactly zero are dropped (:507-511);
    np.round is round-half-even; 'mae' is 1 - mean(real == round(E[r])) (sic)."""
    pred = np.argmax(rat, axis=1)
    keep = rat.sum(axis=1) != 0
    if not keep.all():
        pred, real, rat = pred[keep], real[keep], rat[keep]
    n = len(pred)
    gap = np.abs(pred - real)
    expect = rat @ np.asarray(levels)
    return {
        "accuracy": (gap == 0).astype(int).sum() / n,
        "one_off_accuracy": (gap <= 1).astype(int).sum() / n,
        "mae": 1 - (real == np.round(expect)).astype(int).sum() / n,
        "s2": gap.sum(),
        "s2pond": np.abs(expect - real).sum(),
    }


def choose_best(rats, real, levels):
    """First index of the maximum accuracy, src/mmsbm.py:474-478."""
    acc = [prediction_stats(r, real, levels)["accuracy"] for r in rats]
    return acc.index(max(acc))


# --------------------------------------------------------------------------
# a11  string-rank encoding
# --------------------------------------------------------------------------
def encode_column(values):
    """str() every cell, rank by sorted(set(strings)) -- src/data_handler.py:27-44."""
    strs = [str(v) for v in values]
    table = {s: k for k, s in enumerate(sorted(set(strs)))}
    return np.array([table[s] for s in strs], dtype=np.int64), table


def encode_train(users, items, ratings):
    """int64 [N,3] plus the three dictionaries (src/data_handler.py:46-61)."""
    cu, du = encode_column(users)
    ci, di = encode_column(items)
    cr, dr = encode_column(ratings)
    return np.stack([cu, ci, cr], axis=1), (du, di, dr)


def encode_test(users, items, ratings, dicts):
    """Drop rows with an id unseen in training, column by column in the order
    users, items, ratings, then encode (src/data_handler.py:63-71,109-141)."""
    cols = [[str(v) for v in users], [str(v) for v in items], [str(v) for v in ratings]]
    keep = np.ones(len(cols[0]), dtype=bool)
    for col, table in zip(cols, dicts):
        keep &= np.array([s in table for s in col], dtype=bool)
    out = np.empty((int(keep.sum()), 3), dtype=np.int64)
    for c, (col, table) in enumerate(zip(cols, dicts)):
        out[:, c] = [table[s] for s, k in zip(col, keep) if k]
    return out, keep
