"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE (oracle/_ref/libnavref_*.so,
compiled from /root/reference by oracle/build_ref.sh).  The reference ships no golden vectors
(SURVEY section 4), so these are the pinned known answers for the front-end path.  Run from the
repo root in the container that has /root/reference:  python tests/golden/make_golden.py
"""
import hashlib
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import big_stack  # noqa: E402
from oracle_lib import RefLib  # noqa: E402

synth = importlib.import_module("nav-slam_b200.synth")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small(shape):
    r, c = shape
    ref = RefLib(r, c)
    out = {}
    if shape == (8, 8):
        depth = np.stack([synth.l5_depth_frame(f) for f in range(5)])
        clouds = np.stack([ref.convert(d) for d in depth])
        out["depth"] = depth
    else:
        clouds = np.stack([synth.room_frame(r, c, f, invalid_frac=0.02 if f == 2 else 0.0) for f in range(5)])
    out["clouds"] = clouds
    out["labels"] = np.stack([ref.extract_feature(cl) for cl in clouds])
    out["flat_row0"] = ref.flatten(clouds[1][0], out["labels"][1][0])
    # whole step through the reference's init_slam / slam_localization / slam_mapping
    attr = ref.new_attr()
    pos = np.array([10.0, -20.0, 5.0, 1.0, -2.0, 30.0])
    big_stack(ref.init_slam, attr, pos, clouds[0])
    poses, errors, globals_ = [pos], [0.0], [ref.attr_global(attr, 0).copy()]
    last = pos
    for f in range(1, 5):
        pred = last + np.array([45.0, 3.0, -1.0, 0.0, 0.0, 0.0])
        p = big_stack(ref.slam_localization, attr, clouds[f], pred, last)
        errors.append(ref.attr_error(attr))
        big_stack(ref.slam_mapping, attr, p, clouds[f])
        globals_.append(ref.attr_global(attr, f).copy())
        poses.append(p)
        last = p
    out["poses"] = np.stack(poses)
    out["errors"] = np.array(errors)
    out["globals"] = np.stack(globals_)
    # kd-tree: permutation the reference leaves in the caller's array, pre-order dump, NN answers
    pts = synth.map_points(300, seed=300 + r)
    q = synth.map_queries(pts, 64, seed=301 + r)
    h, perm = ref.tree_build(pts)
    pre, depth_ = ref.tree_preorder(h, 300)
    nn_pt, nn_dist, _ = ref.nn_batch(h, q)
    ref.tree_free(h)
    out.update(kd_pts=pts, kd_q=q, kd_perm=perm, kd_preorder=pre, kd_depth=depth_, kd_nn_pt=nn_pt, kd_nn_dist=nn_dist)
    np.savez_compressed(os.path.join(HERE, f"ref_{r}x{c}.npz"), **out)


def big(shape, cfg_kw):
    r, c = shape
    ref = RefLib(r, c)
    out = {}
    clouds = [synth.room_frame(r, c, f, **cfg_kw) for f in range(3)]
    labels = [ref.extract_feature(cl) for cl in clouds]
    out["label_sha256"] = np.array([sha(x) for x in labels])
    out["label_count"] = np.array([int(x.sum()) for x in labels])
    attr = ref.new_attr()
    pos = np.zeros(6)
    big_stack(ref.init_slam, attr, pos, clouds[0])
    poses, errors, gsha = [pos], [0.0], [sha(ref.attr_global(attr, 0))]
    last = pos
    for f in range(1, 3):
        pred = last + np.array([48.0, 1.0, 0.0, 0.0, 0.0, 0.0])
        p = big_stack(ref.slam_localization, attr, clouds[f], pred, last)
        errors.append(ref.attr_error(attr))
        big_stack(ref.slam_mapping, attr, p, clouds[f])
        gsha.append(sha(ref.attr_global(attr, f)))
        poses.append(p)
        last = p
    out["poses"] = np.stack(poses)
    out["errors"] = np.array(errors)
    out["global_sha256"] = np.array(gsha)
    np.savez_compressed(os.path.join(HERE, f"ref_{r}x{c}.npz"), **out)


if __name__ == "__main__":
    small((8, 8))
    small((5, 33))
    big((16, 1800), dict(cfg=2, elev=(-15, 15), integer_mm=True))
    big((64, 2048), dict())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
