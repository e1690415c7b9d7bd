#!/usr/bin/env python
"""Config 2 (16x1800 integer-millimetre L9 data) tie census: how often the nearest map point of a query is
not unique, and what the choice costs.  The reference returns the first tied point its depth-first search
meets (utils/kdtree.c:116-121), this library the tied point of lowest index (north_star's contract); on fp64
data (configs 1, 3, 4) ties do not occur and the two agree bit for bit.

CPU only: both rules are run by the oracle (tie_mode 0 = the reference's tree and visiting order, pinned
against the compiled reference; tie_mode 1 = lowest index, which the CUDA path reproduces bit for bit --
tests/test_gpu_parity.py).  Per frame: labelled queries, queries whose minimum (sqrt-rounded) distance is
attained by more than one map point of their row, queries for which the two rules pick different points,
correspondences after the dedupe, pose difference, and the difference of the mapped cloud as the CSV prints
it (%.2f).  usage: tie_census.py [n_frames] > profiles/r2_config2_ties.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def census(oracle, frames, step):
    """frames [n, R, C, 3]; returns a list of per-frame dicts (frame 1 .. n-1).  Both rules are fed the
    reference rule's pose as `last`, so every frame compares the two on identical inputs."""
    n, R, C_ = frames.shape[:3]
    ref, can = oracle.slam(R, C_, 0), oracle.slam(R, C_, 1)
    pos0 = np.zeros(6)
    ref.init(pos0, frames[0])
    can.init(pos0, frames[0])
    last, rows = pos0, []
    prev_global = oracle.transform(frames[0].reshape(-1, 3), pos0).reshape(R, C_, 3)
    prev_feat = oracle.extract_feature(frames[0])
    for f in range(1, n):
        pred = last + step
        p_ref, corr_ref, err_ref, _ = ref.localize(frames[f], pred, last)
        p_can, corr_can, err_can, _ = can.localize(frames[f], pred, last)
        feat = oracle.extract_feature(frames[f])
        q_all = oracle.shift(oracle.transform(frames[f].reshape(-1, 3), pred), pred[:3] - last[:3]).reshape(R, C_, 3)
        tied = queries = 0
        for r in range(R):
            pts = np.ascontiguousarray(prev_global[r][prev_feat[r] == 1])
            q = np.ascontiguousarray(q_all[r][feat[r] == 1])
            queries += q.shape[0]
            if pts.shape[0] and q.shape[0]:
                tied += int((oracle.nn_tie_count(pts, q) > 1).sum())
        same_corr = corr_ref.shape == corr_can.shape and np.array_equal(corr_ref, corr_can)
        # entries (query point, matched map point, distance) of one list that the other does not have
        set_ref = {r.tobytes() for r in np.ascontiguousarray(corr_ref)}
        set_can = {r.tobytes() for r in np.ascontiguousarray(corr_can)}
        differing = max(len(set_ref - set_can), len(set_can - set_ref))
        g_ref = ref.map(p_ref, frames[f])
        g_can = can.map(p_ref, frames[f])          # same pose: the maps stay comparable frame after frame
        csv_delta = float(np.abs(np.round(oracle.transform(frames[f].reshape(-1, 3), p_ref), 2)
                                 - np.round(oracle.transform(frames[f].reshape(-1, 3), p_can), 2)).max())
        rows.append({"frame": f, "queries": queries, "tied_queries": tied, "correspondences": int(corr_ref.shape[0]),
                     "correspondences_with_a_different_point": differing, "same_correspondence_list": bool(same_corr),
                     "pose_delta_max_mm": float(np.abs(p_ref - p_can)[:3].max()), "pose_equal_bits": bool(np.array_equal(p_ref, p_can)),
                     "rms_delta_mm": float(abs(err_ref - err_can)), "csv_delta_max": csv_delta,
                     "maps_equal": bool(np.array_equal(g_ref, g_can))})
        prev_global, prev_feat, last = g_ref, feat, p_ref
    ref.close()
    can.close()
    return rows


if __name__ == "__main__":
    import importlib

    from oracle_lib import Oracle
    nav = importlib.import_module("nav-slam_b200")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    frames = nav.synth.l9_sequence(n)
    rows = census(Oracle(), frames, np.array([48.0, 1.0, 0.0, 0.0, 0.0, 0.0]))
    print(f"# config 2 tie census: {n} frames of the synthetic 16x1800 integer-mm L9 sequence (nav-slam_b200/synth.py)")
    print("# frame queries tied_queries corr corr_entries_not_shared pose_delta_max_mm pose_bits_equal csv_delta_max(%.2f)")
    for r in rows:
        print(f"{r['frame']:5d} {r['queries']:7d} {r['tied_queries']:7d} {r['correspondences']:6d} "
              f"{r['correspondences_with_a_different_point']:6d} {r['pose_delta_max_mm']:12.3e} "
              f"{str(r['pose_equal_bits']):5s} {r['csv_delta_max']:.2f}")
    worst = max(r["pose_delta_max_mm"] for r in rows)
    print(f"# max pose delta {worst:.3e} mm, max CSV delta {max(r['csv_delta_max'] for r in rows):.2f}; "
          f"frames without a differing correspondence: {sum(r['same_correspondence_list'] for r in rows)} "
          f"(pose bits equal in all of them: {all(r['pose_equal_bits'] for r in rows if r['same_correspondence_list'])})")
