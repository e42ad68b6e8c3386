"""numpy model of knn2_tc_finish_kernel's selection logic (csrc/knn2_tc_kernel.cu): the distance
kernel keeps, per query, the maximum dot of every bucket of 16 train rows (dot = bits - 2 *
hamming, so a bucket's maximum dot is its exact minimum distance) and hands the finish kernel
the best two buckets as keys  max_dot << 20 | (2^20 - 1 - bucket).  The finish kernel
  1. gives an early verdict from the two keys alone: with d0 the best bucket's minimum distance
     and d2b the second-best bucket's, a query with !(d0 < ratio * d2b) cannot pass the ratio
     test of src/slam_frontend.cc:529-536, whatever the rows of its best bucket are;
  2. rescans only the best bucket of the other queries (exact distances, lowest train index on
     ties) and takes the second neighbour's distance as min(second smallest of that bucket, d2b).
No GPU needed: this checks the argument itself - against the oracle's brute-force GetMatches - on
tie-heavy and ragged inputs, so that the kernel's parity tests are not the only thing holding it."""
import numpy as np
import pytest

import synth
from oracle import native, restate

BUCKET = 16
RATIO = restate.NN_MATCH_RATIO


def hamming_matrix(Q, T):
    x = Q[:, None, :] ^ T[None, :, :]
    return np.unpackbits(x, axis=2).sum(axis=2).astype(np.int64)


def finish_model(Q, T, ratio):
    """-> (matches like GetMatches: (queryIdx, trainIdx, distance) in query order, rescans done)"""
    nq, nt = len(Q), len(T)
    if nq == 0 or nt == 0:
        return [], 0
    bits = 8 * Q.shape[1]
    D = hamming_matrix(Q, T)
    nb = (nt + BUCKET - 1) // BUCKET
    pad = np.full((nq, nb * BUCKET), 10 ** 6, np.int64)
    pad[:, :nt] = D
    bucket_min = pad.reshape(nq, nb, BUCKET).min(axis=2)
    # keys: larger = smaller distance, lower bucket on ties
    keys = (bits - 2 * bucket_min) * (1 << 20) + ((1 << 20) - 1 - np.arange(nb))[None, :]
    order = np.argsort(-keys, axis=1, kind="stable")
    out, rescans = [], 0
    for q in range(nq):
        b1 = order[q, 0]
        d0b = int(bucket_min[q, b1])
        have2 = nb > 1
        d2b = int(bucket_min[q, order[q, 1]]) if have2 else None
        if have2 and ratio >= 0.0 and not (float(d0b) < ratio * float(d2b)):
            continue                                   # early verdict: fails, no rescan
        rescans += 1
        rows = np.arange(b1 * BUCKET, min(nt, (b1 + 1) * BUCKET))
        d = D[q, rows]
        o = np.lexsort((rows, d))                      # (distance, trainIdx)
        cand = [(int(d[i]), int(rows[i])) for i in o[:2]]
        if have2:
            cand.append((d2b, -1))                     # index never read
            cand.sort(key=lambda t: t[0])              # only the distance of the second matters
        if len(cand) < 2:
            continue
        (d0, i0), (d1, _) = cand[0], cand[1]
        assert d0 == d0b
        if float(d0) < ratio * float(d1):
            out.append((q, i0, d0))
    return out, rescans


def check(Q, T, ratio=RATIO):
    got, rescans = finish_model(Q, T, ratio)
    exp = native.get_matches(Q, T, ratio)
    assert [(int(m["queryIdx"]), int(m["trainIdx"]), int(m["distance"])) for m in exp] == got
    return rescans


@pytest.mark.parametrize("nq,nt,seed", [(300, 300, 1), (257, 129, 2), (64, 17, 3), (50, 16, 4), (40, 2, 5), (33, 1, 6)])
def test_early_verdict_and_one_bucket_rescan_equal_brute_force(nq, nt, seed):
    Q, T = synth.descriptor_pair(nq, nt, seed=seed)
    rescans = check(Q, T)
    assert rescans <= nq


def test_ties_and_duplicates():
    rng = np.random.default_rng(7)
    T = rng.integers(0, 256, (96, 32), dtype=np.uint8)
    T[40] = T[3]                       # the same row in two buckets
    T[41] = T[3]
    T[17] = T[16]                      # and twice inside one bucket
    Q = T[rng.integers(0, 96, 200)].copy()
    flip = rng.integers(0, 256, 200)
    Q[np.arange(200), flip // 8] ^= (1 << (flip % 8)).astype(np.uint8)
    Q[:20] = T[3]                      # exact duplicates of a duplicated row
    check(Q, T)
    check(Q, T, ratio=1.0)
    check(Q, T, ratio=0.0)


@pytest.mark.parametrize("width", [61, 64])
def test_wide_rows(width):
    Q, T = synth.descriptor_pair(150, 170, seed=9)
    rng = np.random.default_rng(10)
    Qw = np.concatenate([Q, rng.integers(0, 256, (150, width - 32), dtype=np.uint8)], axis=1)
    Tw = np.concatenate([T, rng.integers(0, 256, (170, width - 32), dtype=np.uint8)], axis=1)
    Tw[:100, 32:] = Qw[:100, 32:]      # keep the planted near-duplicates near
    check(Qw, Tw)


def test_most_uniform_queries_are_rejected_without_a_rescan():
    """What the early verdict buys: uniformly random descriptors never pass Lowe's test (SURVEY
    8c fact 5) and almost all of them are rejected from the two bucket keys alone."""
    rng = np.random.default_rng(11)
    Q = rng.integers(0, 256, (400, 32), dtype=np.uint8)
    T = rng.integers(0, 256, (800, 32), dtype=np.uint8)
    rescans = check(Q, T)
    assert rescans < 20


def test_look_back_words_fit_the_problem_slots():
    """FinishArgs::flags is indexed in 32-query units: run_knn reserves round_up(nq, 128) / 32
    words per problem (KnnProblem::qb0), a block of QPB queries publishes into word
    block * (QPB / 32) - whatever QPB the launch picks, the last block's word is inside."""
    for qpb in (64, 128, 256):
        for nq in list(range(1, 1200)) + [4999, 5000, 5001, 20000, 24576]:
            words = ((max(nq, 1) + 127) // 128 * 128) // 32
            nqb = (nq + qpb - 1) // qpb
            assert (nqb - 1) * (qpb // 32) < words, (qpb, nq)
