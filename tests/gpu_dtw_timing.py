"""Latency / throughput of the DTW family at the 25-room training shape (B = 16 sequences, 255 tree
nodes x 200 frames of 3x32x32) on the B200, with the CPU oracle port timed beside it.  Measurement aid, not a test:
    python tests/gpu_dtw_timing.py > profiles/<round>_dtw.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dtw_oracle as D  # noqa: E402
from video_gcp_b200 import dtw  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    eng = dtw.get_engine("cuda:0")
    r = np.random.default_rng(0)
    B, n, m, dim = 16, 255, 200, 3072
    x = torch.from_numpy(r.uniform(-1, 1, size=(B, n, dim)).astype(np.float32)).cuda()
    y = torch.from_numpy(r.uniform(-1, 1, size=(B, m, dim)).astype(np.float32)).cuda()
    end = torch.from_numpy(r.integers(10, m, size=(B,))).cuda()
    t = timed(lambda: eng.cdist_mean(x, y))
    pair_elems = B * n * m * dim
    print("cdist_mean  B=%d %dx%d dim %d: %.3f ms  (%.2f T pair-elements/s = %.1f TFLOP/s fp32 at 3 flop/element; "
          "input %.1f MB -> %.0f GB/s)" % (B, n, m, dim, t, pair_elems / t / 1e9, 3 * pair_elems / t / 1e9,
                                          (x.numel() + y.numel()) * 4 / 1e6, (x.numel() + y.numel()) * 4 / t / 1e6))
    cost = eng.cdist_mean(x, y).clone()
    t = timed(lambda: eng.soft_dtw(cost, 1.0, end, want_bf=True))
    print("soft_dtw    B=%d %dx%d (sweep + weights + row-sum check + normalise/bf): %.3f ms  (%.1f M cells/s per direction)"
          % (B, n, m, t, B * n * m / t / 1e3))
    t = timed(lambda: eng.dtw(cost, end))
    print("dtw         B=%d %dx%d (wavefront + traceback + matches): %.3f ms" % (B, n, m, t))
    xc, yc, cc, ec = x.cpu(), y.cpu(), cost.cpu(), end.cpu().numpy()
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter(); D.batch_cdist_mean(xc, yc); t1 = time.perf_counter()
    print("CPU port (%d threads) cdist_mean: %.1f ms" % (os.cpu_count(), (t1 - t0) * 1e3))
    t0 = time.perf_counter(); D.binding_weights(cc, 1.0, ec); t1 = time.perf_counter()
    print("CPU port soft_dtw + normalise: %.1f ms" % ((t1 - t0) * 1e3))
    t0 = time.perf_counter(); D.batched_dtw(cc[:2].numpy().astype(np.float64), ec[:2]); t1 = time.perf_counter()
    print("CPU port dtw (python loops, 2 of %d sequences): %.1f ms -> %.1f ms for the batch" % (B, (t1 - t0) * 1e3, (t1 - t0) * 1e3 * B / 2))


if __name__ == "__main__":
    main()
