"""Pins oracle/dtw_oracle.py (DTW family) against fixtures produced by the unmodified reference
(oracle/make_golden_dtw.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dtw_oracle as D


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "dtw_family.npz"))


def full_cost(g):
    return np.random.default_rng(int(g["full_seed"])).uniform(0.0, 1.3, size=(2, 255, 200)).astype(np.float32)


def test_gak_tables(g):
    C = -g["soft_cost"].astype(np.float64)
    for b in range(3):
        np.testing.assert_allclose(D.gak_table(C[b], 0), g["soft_fwd"][b], rtol=1e-13, atol=1e-13)
        begin = 8 - int(g["soft_end"][b]) - 1
        np.testing.assert_allclose(D.gak_table(C[b, ::-1, ::-1], begin), g["soft_bwd_flipped"][b], rtol=1e-13, atol=1e-13)


def test_gak_column_wraparound_and_single_column(g):
    for b in range(2):
        got = D.gak_table(g["wrap_C"][b], int(g["wrap_begin"][b]))
        np.testing.assert_allclose(got, g["wrap_D"][b], rtol=1e-13, atol=1e-13)
    assert np.isfinite(g["wrap_D"][0][1, 0])          # the leak through column -1 is really there in the reference
    np.testing.assert_allclose(D.gak_table(g["onecol_C"][0], 0), g["onecol_D"][0], rtol=1e-13, atol=1e-13)


def test_soft_dtw_small(g):
    w, _, _ = D.soft_dtw(g["soft_cost"], g["soft_end"])
    np.testing.assert_allclose(w, g["soft_w"], rtol=1e-6, atol=1e-12)
    w, _, _ = D.soft_dtw(g["soft_cost"])
    np.testing.assert_allclose(w, g["soft_w_noend"], rtol=1e-6, atol=1e-12)
    w, _, _ = D.soft_dtw(g["sq_cost"])
    np.testing.assert_allclose(w, g["sq_w"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(w, np.broadcast_to(np.eye(7, dtype=np.float32), (2, 7, 7)), atol=1e-6)


def test_soft_dtw_full_size(g):
    w, _, _ = D.soft_dtw(full_cost(g), g["full_end"])
    np.testing.assert_allclose(w[:, ::16], g["full_w_rows"], rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(w.astype(np.float64).sum(1), g["full_w_sum_nodes"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(w.astype(np.float64).sum(2), g["full_w_sum_frames"], rtol=1e-6)
    np.testing.assert_array_equal(w.argmax(2), g["full_argmax_frame"])
    np.testing.assert_allclose(w.sum(2), 1.0, atol=1e-4)                 # every node binds to exactly one frame
    assert (w[1, :, 58:] == 0).all()                                     # frames past end_ind never bind


def test_binding_weights(g):
    cost = D.batch_cdist_mean(torch.from_numpy(g["getw_imgs"]), torch.from_numpy(g["getw_traj"]))
    np.testing.assert_allclose(cost.numpy(), g["getw_cost"], rtol=1e-6, atol=1e-7)
    w = D.binding_weights(cost, float(g["getw_temp"]), g["getw_end"])
    np.testing.assert_allclose(w.numpy(), g["getw_w"], rtol=1e-5, atol=1e-9)
    np.testing.assert_array_equal(D.df_to_bf_order(15), [7, 3, 11, 1, 5, 9, 13, 0, 2, 4, 6, 8, 10, 12, 14])


@pytest.mark.parametrize("case", ["dtw", "tie"])
def test_basic_dtw_bit_exact(g, case):
    d, acc, (p, q) = D.basic_dtw(g[case + "_cost"])
    assert d == float(g[case + "_dist"])
    np.testing.assert_array_equal(acc, g[case + "_acc"])
    np.testing.assert_array_equal(p, g[case + "_p"])
    np.testing.assert_array_equal(q, g[case + "_q"])


def test_batched_dtw(g):
    dist, acc, paths, lengths = D.batched_dtw(g["bat_cost"].astype(np.float64), g["bat_end"])
    np.testing.assert_array_equal(dist, g["bat_dist"])
    np.testing.assert_array_equal(acc, g["bat_acc"])
    P, Q = D.stack_batched_paths(paths)
    np.testing.assert_array_equal(P, g["bat_P"])
    np.testing.assert_array_equal(Q, g["bat_Q"])
    np.testing.assert_array_equal(lengths, g["bat_len"])


def test_single_matches(g):
    inds, (p, q), _, _ = D.single_matches(torch.from_numpy(g["match_est"]), torch.from_numpy(g["match_tgt"]))
    np.testing.assert_array_equal(inds, g["match_inds"])
    np.testing.assert_array_equal(p, g["match_p"])
    np.testing.assert_array_equal(q, g["match_q"])


def test_min_cumsum_matches_the_reference_compiled_extension():
    """The oracle's accumulated-cost table against the reference's own native code: gcp/evaluation/cutils.pyx compiled by
    oracle/build_ref_cutils.py (c_dtw's inner loop, dtw_utils.py:99-116).  Bit-exact on float64, incl. ragged shapes, ties,
    a single row / column and infinities in the cost."""
    from oracle import build_ref_cutils
    cutils = build_ref_cutils.load()
    if cutils is None:
        pytest.skip("oracle/_ref/cutils*.so not built (python -m oracle.build_ref_cutils needs /root/reference)")
    r = np.random.default_rng(3)
    cases = [r.uniform(0, 5, size=s) for s in ((1, 1), (1, 9), (9, 1), (7, 13), (40, 31), (200, 200))]
    cases.append(r.integers(0, 3, size=(25, 25)).astype(np.float64))        # ties
    inf_case = r.uniform(0, 5, size=(12, 12))
    inf_case[3, :5] = np.inf
    cases.append(inf_case)
    for C in cases:
        rr, cc = C.shape
        T = np.zeros((rr + 1, cc + 1))
        T[0, 1:] = np.inf
        T[1:, 0] = np.inf
        T[1:, 1:] = C
        cutils.min_cumsum(T)                       # in place, as c_dtw calls it
        assert np.array_equal(T, D.min_cumsum(C)), C.shape
