"""GPU parity for the DTW family: csrc/dtw_kernels.cuh through the C ABI against oracle/dtw_oracle.py
and the fixtures recorded from the unmodified reference (tests/golden/dtw_family.npz).  Run with -m gpu on the B200.

Tolerances: float64 soft-DTW tables 1e-11 relative (device exp/log vs libm, <= 1 ulp each over <= 255 chained rows);
weights w (float32 outputs) 2e-6 relative; cost matrices 2e-6 of the matrix scale (the device sums squared differences,
the reference expands |x|^2 + |y|^2 - 2xy in fp32); accumulated DTW tables, distances, warping paths and per-frame
matches bit-exact (float64 additions and comparisons in the reference's order).
"""
import os

import numpy as np
import pytest
import torch

from oracle import dtw_oracle as D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "dtw_family.npz"))


@pytest.fixture(scope="module")
def eng():
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    from video_gcp_b200 import dtw
    return dtw.get_engine("cuda:0")


def full_cost(g):
    return np.random.default_rng(int(g["full_seed"])).uniform(0.0, 1.3, size=(2, 255, 200)).astype(np.float32)


def _np(t):
    return t.detach().cpu().numpy()


def test_soft_dtw_tables_and_weights_small(g, eng):
    out = eng.soft_dtw(torch.from_numpy(g["soft_cost"]), 1.0, g["soft_end"], want_tables=True)
    np.testing.assert_allclose(_np(out["forward"]), g["soft_fwd"], rtol=1e-11, atol=1e-11)
    bwd_ref = g["soft_bwd_flipped"][:, ::-1, ::-1]
    np.testing.assert_allclose(_np(out["backward"]), bwd_ref, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(_np(out["w"]), g["soft_w"], rtol=2e-6, atol=1e-12)
    assert abs(float(out["rowsum_max"]) - 1) < 1e-4
    out = eng.soft_dtw(torch.from_numpy(g["soft_cost"]), 1.0, None)
    np.testing.assert_allclose(_np(out["w"]), g["soft_w_noend"], rtol=2e-6, atol=1e-12)
    out = eng.soft_dtw(torch.from_numpy(g["sq_cost"]), 1.0, None)
    np.testing.assert_allclose(_np(out["w"]), g["sq_w"], rtol=2e-6, atol=1e-12)


def test_soft_dtw_wraparound_cases(g, eng):
    """The reference's column -1 wrap-around: checked on the backward table (flipped problem beginning in the last
    column <=> end_ind = 0 on the un-flipped one) and on a single-column problem."""
    cost = (-g["wrap_C"]).astype(np.float32)                 # [2,5,3]; as float32 so that the device sees the same values
    ref = [D.gak_table(-cost[b].astype(np.float64), int(g["wrap_begin"][b])) for b in range(2)]
    # a flipped cost with end = c - begin - 1 makes the backward pass run exactly this problem
    flipped = np.ascontiguousarray(cost[:, ::-1, ::-1])
    end = 3 - g["wrap_begin"] - 1
    out = eng.soft_dtw(torch.from_numpy(flipped), 1.0, end, want_tables=True)
    got = _np(out["backward"])[:, ::-1, ::-1]
    np.testing.assert_allclose(got, np.stack(ref), rtol=1e-12, atol=1e-12)
    assert np.isfinite(got[0, 1, 0])
    one = (-g["onecol_C"]).astype(np.float32)
    out = eng.soft_dtw(torch.from_numpy(one), 1.0, None, want_tables=True)
    np.testing.assert_allclose(_np(out["forward"])[0], D.gak_table(-one[0].astype(np.float64), 0), rtol=1e-12)


def test_soft_dtw_full_size(g, eng):
    fc = full_cost(g)
    out = eng.soft_dtw(torch.from_numpy(fc), 1.0, g["full_end"], want_tables=True)
    w = _np(out["w"])
    np.testing.assert_allclose(w[:, ::16], g["full_w_rows"], rtol=4e-6, atol=1e-12)
    np.testing.assert_allclose(w.astype(np.float64).sum(1), g["full_w_sum_nodes"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(w.astype(np.float64).sum(2), g["full_w_sum_frames"], rtol=2e-6)
    np.testing.assert_array_equal(w.argmax(2), g["full_argmax_frame"])
    w_ref, fwd, bwd = D.soft_dtw(fc, g["full_end"])
    np.testing.assert_allclose(_np(out["forward"]), fwd, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(_np(out["backward"]), bwd, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(w, w_ref, rtol=4e-6, atol=1e-12)


def test_temperature_is_the_reference_division(g, eng):
    """cost / temp in fp32 before the float64 sweep (adaptive.py:51): same bits as dividing on the host."""
    cost = torch.from_numpy(g["soft_cost"])
    a = _np(eng.soft_dtw(cost, 0.37, g["soft_end"])["w"]).copy()
    b = _np(eng.soft_dtw(cost / torch.full((1,), 0.37), 1.0, g["soft_end"])["w"])
    np.testing.assert_array_equal(a, b)


def test_binding_weights_chain(g, eng):
    from video_gcp_b200 import dtw
    cost = dtw.batch_cdist(torch.from_numpy(g["getw_imgs"]).cuda(), torch.from_numpy(g["getw_traj"]).cuda(), 'mean')
    scale = float(np.abs(g["getw_cost"]).max())
    assert np.abs(_np(cost) - g["getw_cost"]).max() < 2e-6 * scale
    # same cost matrix in -> the reference's weights out
    w = dtw.get_w(torch.from_numpy(g["getw_cost"]).cuda(), float(g["getw_temp"]), g["getw_end"])
    np.testing.assert_allclose(_np(w), g["getw_w"], rtol=1e-5, atol=1e-9)
    # the whole chain (device cost matrix): exp() amplifies the 1e-6 cost differences by cost / temp
    w2 = dtw.get_w(cost, float(g["getw_temp"]), g["getw_end"])
    np.testing.assert_allclose(_np(w2), g["getw_w"], rtol=1e-4, atol=1e-8)


def test_cdist_full_size_and_ragged_tiles(eng):
    r = np.random.default_rng(5)
    for (B, n, m, dim) in ((2, 255, 200, 3072), (3, 70, 33, 128), (1, 1, 1, 4)):
        x = torch.from_numpy(r.uniform(-1, 1, size=(B, n, dim)).astype(np.float32))
        y = torch.from_numpy(r.uniform(-1, 1, size=(B, m, dim)).astype(np.float32))
        y[0, 0] = x[0, 0]                                      # identical vectors: distance exactly 0 on the device
        got = _np(eng.cdist_mean(x, y))
        ref = ((x.double()[:, :, None] - y.double()[:, None]) ** 2).mean(-1).numpy() if n * m * dim < 4e7 else None
        orc = D.batch_cdist_mean(x, y).numpy()
        assert np.abs(got - orc).max() <= 2e-6 * np.abs(orc).max(), (B, n, m, dim)
        if ref is not None:
            assert np.abs(got - ref).max() <= 5e-7 * np.abs(ref).max()
        assert got[0, 0, 0] == 0.0


@pytest.mark.parametrize("case", ["dtw", "tie"])
def test_c_dtw_bit_exact(g, case):
    from video_gcp_b200 import dtw
    d, acc, (p, q) = dtw.c_dtw(g[case + "_cost"])
    assert d == float(g[case + "_dist"])
    np.testing.assert_array_equal(acc, g[case + "_acc"])
    np.testing.assert_array_equal(p, g[case + "_p"])
    np.testing.assert_array_equal(q, g[case + "_q"])


def test_batched_dtw_matches_reference_including_quirks(g):
    from video_gcp_b200 import dtw
    end = g["bat_end"].copy()
    dist, acc, (P, Q), lengths = dtw.batched_dtw(g["bat_cost"].astype(np.float64), end)
    np.testing.assert_array_equal(dist, g["bat_dist"])
    np.testing.assert_array_equal(acc, g["bat_acc"])
    np.testing.assert_array_equal(P, g["bat_P"])
    np.testing.assert_array_equal(Q, g["bat_Q"])
    np.testing.assert_array_equal(lengths, g["bat_len"])
    assert (end == 0).all()                                    # the reference zeroes the caller's array


def test_dtw_full_size_against_oracle(eng):
    r = np.random.default_rng(11)
    C = r.uniform(0, 1, size=(4, 255, 200)).astype(np.float32)
    end = np.array([199, 57, 0, 120])
    out = eng.dtw(torch.from_numpy(C), end)
    dist_o, acc_o, paths_o, _ = D.batched_dtw(C.astype(np.float64), end)
    np.testing.assert_array_equal(_np(out["acc"])[:, 1:, 1:], acc_o)
    for b in range(4):
        n = int(out["path_len"][b])
        np.testing.assert_array_equal(_np(out["path_p"])[b, -n:], paths_o[b][0])
        np.testing.assert_array_equal(_np(out["path_q"])[b, -n:], paths_o[b][1])
        assert (_np(out["path_p"])[b, :-n] == 0).all() and (_np(out["path_q"])[b, :-n] == 0).all()
        acc = acc_o[b]
        match = np.full_like(acc, np.inf)
        match[paths_o[b][0], paths_o[b][1]] = acc[paths_o[b][0], paths_o[b][1]]
        np.testing.assert_array_equal(_np(out["match_inds"])[b], np.argmin(match, axis=0))
        assert float(out["dist"][b]) == acc[-1, end[b]] / (255 + end[b] + 1)


def test_single_matches(g):
    from video_gcp_b200 import dtw
    gen, mo = dtw.DTWEvalBinding.get_single_matches(torch.from_numpy(g["match_tgt"]), torch.from_numpy(g["match_est"]))
    np.testing.assert_array_equal(mo.matching_path[0], g["match_p"])
    np.testing.assert_array_equal(mo.matching_path[1], g["match_q"])
    np.testing.assert_array_equal(_np(gen), g["match_est"][g["match_inds"]])


def test_rejects_bad_arguments(eng):
    from video_gcp_b200 import _C
    with pytest.raises(_C.GcpB200Error):
        eng.soft_dtw(torch.zeros(1, 4, 9), 1.0, None)           # fewer nodes than frames (reference asserts r >= c)
    with pytest.raises(_C.GcpB200Error):
        eng.soft_dtw(torch.zeros(1, 6, 4), 1.0, None, want_bf=True)   # breadth-first order needs 2^d - 1 nodes
    with pytest.raises(_C.GcpB200Error):
        eng.cdist_mean(torch.zeros(1, 2, 6), torch.zeros(1, 2, 6))
    out = eng.soft_dtw(torch.zeros(1, 4, 3), 1.0, np.array([5]))      # end index outside the sequence: NaN, never garbage
    assert torch.isnan(out["w"]).all()
    from video_gcp_b200 import dtw
    with pytest.raises(FloatingPointError):
        dtw.soft_dtw(torch.zeros(1, 4, 3).cuda(), np.array([5]))
