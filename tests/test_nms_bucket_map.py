"""CPU: the invariants the NMS kernel's bucket rank sort rests on (`csrc/nms_match.cu::bucket_sort`), restated in numpy
with the kernel's fp32 operations: the key -> bucket map must be non-decreasing in the key (equal keys share a bucket)
for ANY min / max, and bucket start + in-bucket rank of the (key, index) pairs must be the stable descending-score order
`torchvision.ops.nms` consumes its candidates in (`/root/reference/src/running_main_v2.py:817`: ties -> lower index)."""
import numpy as np
import pytest

BUCKETS = 2048


def desc_key(s):
    """nms_match.cu::desc_key: smaller key = higher score, NaN first, -0.0 == +0.0."""
    s = np.asarray(s, np.float32)
    u = s.view(np.uint32).copy()
    u[u == 0x80000000] = 0
    asc = np.where(u & 0x80000000, ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    k = (~asc).astype(np.uint32)
    k[np.isnan(s)] = 0
    return k


def bucket_of(k, kmin, kmax):
    """fp32, one rounding per operation, truncation, clamp -- as the kernel does it."""
    scale = np.float32(BUCKETS) / (np.float32(np.uint32(kmax - kmin)) + np.float32(1.0))
    b = ((k - kmin).astype(np.uint32).astype(np.float32) * np.float32(scale)).astype(np.int64)
    return np.minimum(b, BUCKETS - 1)


def key_sets():
    rng = np.random.default_rng(5)
    yield np.sort(rng.integers(0, 2**32, 20000, dtype=np.uint64).astype(np.uint32))            # whole range
    yield np.sort((np.uint32(0x3f000000) + rng.integers(0, 5000, 20000)).astype(np.uint32))   # narrow band
    yield np.sort(np.concatenate([np.arange(2**24 - 300, 2**24 + 300), np.arange(2**31 - 300, 2**31 + 300),
                                  [0, 1, 2**32 - 2, 2**32 - 1]]).astype(np.uint32))             # fp32 conversion breakpoints
    yield np.full(100, 12345, np.uint32)                                                        # one key
    yield np.sort(desc_key(rng.uniform(0.05, 1.0, 20000).astype(np.float32)))                  # what the benchmark looks like
    yield np.sort(desc_key((np.round(rng.uniform(0, 1, 20000) ** 2 * 64) / 64).astype(np.float32)))


@pytest.mark.parametrize("case", range(6))
def test_bucket_map_is_monotone_and_in_range(case):
    k = list(key_sets())[case]
    b = bucket_of(k, k.min(), k.max())
    assert b.min() >= 0 and b.max() <= BUCKETS - 1
    assert (np.diff(b) >= 0).all(), "a larger key landed in a smaller bucket"
    same = np.diff(k.astype(np.int64)) == 0
    assert (np.diff(b)[same] == 0).all(), "equal keys in different buckets"


@pytest.mark.parametrize("seed", range(4))
def test_bucket_start_plus_in_bucket_rank_is_the_stable_order(seed):
    rng = np.random.default_rng(100 + seed)
    n = [37, 930, 4096, 8400][seed]
    s = rng.uniform(0.05, 1.0, n).astype(np.float32)
    s[rng.integers(0, n, n // 4)] = s[rng.integers(0, n, n // 4)]          # ties
    if seed == 0:
        s[3], s[5], s[7] = np.nan, -0.0, 0.0
    k = desc_key(s)
    pair = (k.astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    b = bucket_of(k, k.min(), k.max())
    count = np.bincount(b, minlength=BUCKETS)
    start = np.concatenate([[0], np.cumsum(count)[:-1]])
    scattered = rng.permutation(n)                                           # the order the atomics happened to run in
    out = np.empty(n, np.int64)
    for bk in np.unique(b):
        members = scattered[b[scattered] == bk]
        for i in members:
            out[start[bk] + int((pair[members] < pair[i]).sum())] = i
    # the reference order: descending score, ties -> lower index, NaN first (torch sorts NaN as the largest value)
    want = np.argsort(pair, kind="stable")
    np.testing.assert_array_equal(out, want)
    finite = ~np.isnan(s[want])
    assert (np.diff(s[want][finite]) <= 0).all()
