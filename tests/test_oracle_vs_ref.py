"""Pins the oracle: every function of oracle/navslam_oracle.c must be bit-identical to the
reference's own code (oracle/_ref/libnavref_<RxC>.so, built from /root/reference by
oracle/build_ref.sh) on seeded inputs.  CPU only."""
import numpy as np
import pytest

from conftest import big_stack
from oracle_lib import RefLib, ref_available

SHAPES = [(8, 8), (5, 33), (16, 1800), (64, 2048)]


def _ref(shape):
    name = f"{shape[0]}x{shape[1]}"
    if not ref_available(name):
        pytest.skip(f"oracle/_ref/libnavref_{name}.so not built (needs /root/reference)")
    return RefLib(*shape)


def _cloud(synth, shape, frame=0, **kw):
    r, c = shape
    if shape == (16, 1800):
        return synth.room_frame(r, c, frame, cfg=2, elev=(-15, 15), integer_mm=True, **kw)
    return synth.room_frame(r, c, frame, **kw)


@pytest.mark.parametrize("shape", SHAPES)
def test_struct_sizes(shape):
    ref = _ref(shape)
    r, c = shape
    assert ref.sizeof_pointcloud == 8 + r * c * 24
    assert ref.sizeof_slam_attr == 100 * ref.sizeof_pointcloud + 8 + r * 8 + 8
    assert ref.lib.refdrv_sizeof_kdnode() == 40
    assert ref.lib.refdrv_sizeof_neighbor_result() == 56


@pytest.mark.parametrize("shape", SHAPES)
def test_extract_feature_bit_exact(oracle, synth, shape):
    ref = _ref(shape)
    for frame, inv in ((0, 0.0), (3, 0.02)):
        cloud = _cloud(synth, shape, frame, invalid_frac=inv)
        a = ref.extract_feature(cloud)
        b = oracle.extract_feature(cloud)
        assert np.array_equal(a, b)
        assert a[:, :2].sum() == 0 and a[:, -2:].sum() == 0
        curv = oracle.curvature(cloud)
        assert np.array_equal(curv > 0.1, a == 1)
    # never writes zeros: pre-set labels survive
    pre = np.full(shape, 7, dtype=np.int32)
    a = ref.extract_feature(cloud, pre.copy())
    b = oracle.extract_feature(cloud, pre.copy())
    assert np.array_equal(a, b) and set(np.unique(a)) <= {1, 7}


def test_extract_feature_special_values(oracle):
    ref = _ref((5, 33))
    rng = np.random.default_rng(7)
    cloud = rng.normal(0, 1000, size=(5, 33, 3))
    cloud[0, 5] = np.nan
    cloud[1, 7] = np.inf
    cloud[2, 10:20] = 0.0          # run of invalid returns
    cloud[3, :] = cloud[3, 0]      # all equal -> avg 0
    cloud[4, 10:14] = 1e-200       # underflowing squares
    assert np.array_equal(ref.extract_feature(cloud), oracle.extract_feature(cloud))


def test_collinear_equal_spacing_is_edge(oracle):
    # equally spaced collinear neighbours give 0.25/2.25 = 1/9 > 0.1
    cloud = np.zeros((1, 16, 3))
    cloud[0, :, 0] = 100.0 * np.arange(16)
    curv = oracle.curvature(cloud)
    assert np.allclose(curv[0, 2:-2], 1.0 / 9.0, rtol=1e-6)
    assert oracle.extract_feature(cloud)[0, 2:-2].all()


@pytest.mark.parametrize("shape", [(8, 8), (5, 33)])
def test_convert_bit_exact(oracle, synth, shape):
    ref = _ref(shape)
    rng = np.random.default_rng(11)
    d = rng.integers(-5, 4000, size=shape).astype(np.int32)
    d[0, 0] = 0
    assert np.array_equal(ref.convert(d), oracle.convert(d))
    d8 = synth.l5_depth_frame(3, *shape)
    assert np.array_equal(ref.convert(d8), oracle.convert(d8))


def test_rotation_and_transform_bit_exact(oracle, synth):
    ref = _ref((5, 33))
    rng = np.random.default_rng(5)
    for _ in range(20):
        pos = np.concatenate([rng.normal(0, 5000, 3), rng.uniform(-180, 180, 3)])
        rad = pos[3:] * np.pi / 180.0
        R_ref = ref.rotation(pos[3] * np.pi / 180.0, pos[4] * np.pi / 180.0, pos[5] * np.pi / 180.0)
        assert np.array_equal(R_ref, oracle.rotation_deg(pos))
        del rad


@pytest.mark.parametrize("shape", SHAPES)
def test_flatten_bit_exact(oracle, synth, shape):
    ref = _ref(shape)
    cloud = _cloud(synth, shape, 1)
    feat = oracle.extract_feature(cloud)
    feat[0, 0] = 1
    feat[-1, -1] = 1
    feat[0, 1] = 2  # only ==1 is selected (src/slam.c:68)
    for r in (0, shape[0] - 1):
        assert np.array_equal(ref.flatten(cloud[r], feat[r]), oracle.flatten(cloud[r], feat[r]))


@pytest.mark.parametrize("kind", ["uniform", "clustered", "integer", "duplicates", "sorted"])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 1000, 20000])
def test_tree_build_and_nn_bit_exact(oracle, synth, kind, n):
    ref = _ref((8, 8))
    rng = np.random.default_rng(n + 13)
    if kind == "uniform":
        pts = synth.map_points(n, seed=n)
    elif kind == "clustered":
        pts = synth.map_points(n, variant="clustered", seed=n)
    elif kind == "integer":
        pts = rng.integers(0, 50, size=(n, 3)).astype(np.float64)  # many equal keys + exact ties
    elif kind == "duplicates":
        pts = np.repeat(rng.normal(0, 100, size=((n + 3) // 4, 3)), 4, axis=0)[:n]
    else:
        pts = np.stack([np.arange(n, dtype=np.float64)] * 3, axis=1)  # Lomuto worst case
        if n > 5000:
            pytest.skip("quadratic quick-select")
    h_ref, perm_ref = ref.tree_build(pts)
    h_or, perm_or = oracle.tree_build(pts)
    assert np.array_equal(perm_ref, perm_or)            # same in-place permutation
    if n:
        a, da = ref.tree_preorder(h_ref, n)
        b, db = oracle.tree_preorder(h_or, n)
        assert np.array_equal(a, b) and np.array_equal(da, db)
    nq = 300
    q = (pts[rng.integers(0, n, size=nq)] + rng.normal(0, 3, size=(nq, 3))) if n else rng.normal(size=(nq, 3))
    if kind == "integer":
        q = np.rint(q)
    p_ref, d_ref, _ = ref.nn_batch(h_ref, q)
    p_or, d_or = oracle.tree_nn(h_or, q)
    assert np.array_equal(d_ref, d_or)
    assert np.array_equal(p_ref, p_or, equal_nan=True)
    # canonical form: same distance always; same point unless an exact distance tie exists
    idx, d_can = oracle.nn_brute(pts, q)
    if n:
        assert np.array_equal(d_can, d_ref)
        ties = oracle.nn_tie_count(pts, q)
        same = np.all(pts[idx] == p_ref, axis=1)
        assert np.all(same | (ties > 1))
        # the reference's answer is always one of the tied candidates
        d_back = np.sqrt(((p_ref - q) ** 2)[:, 0] + ((p_ref - q) ** 2)[:, 1] + ((p_ref - q) ** 2)[:, 2])
        assert np.array_equal(d_back, d_ref)
    else:
        assert np.all(idx == -1) and np.all(np.isinf(d_can)) and np.all(np.isinf(d_ref))
    ref.tree_free(h_ref)
    oracle.tree_free(h_or)


@pytest.mark.parametrize("shape", [(8, 8), (5, 33), (16, 1800)])
def test_whole_step_bit_exact(oracle, synth, shape):
    """init_slam / slam_localization / slam_mapping (reference tie mode) vs the reference."""
    ref = _ref(shape)
    r, c = shape
    frames = 6 if shape != (16, 1800) else 3
    attr = ref.new_attr()
    slam = oracle.slam(r, c, 0)
    if shape == (8, 8):
        clouds = [oracle.convert(synth.l5_depth_frame(f)) for f in range(frames)]
    else:
        clouds = [_cloud(synth, shape, f) for f in range(frames)]
    pos = np.array([10.0, -20.0, 5.0, 1.0, -2.0, 30.0])
    big_stack(ref.init_slam, attr, pos, clouds[0])
    g = slam.init(pos, clouds[0])
    assert np.array_equal(ref.attr_global(attr, 0), g)
    last = pos
    for f in range(1, frames):
        pred = last + np.array([45.0, 3.0, -1.0, 0.0, 0.0, 0.0])
        p_ref = big_stack(ref.slam_localization, attr, clouds[f], pred, last)
        p_or, corr, err, iters = slam.localize(clouds[f], pred, last)
        assert np.array_equal(p_ref, p_or), (f, p_ref, p_or)
        assert ref.attr_error(attr) == err
        big_stack(ref.slam_mapping, attr, p_ref, clouds[f])
        g = slam.map(p_or, clouds[f])
        assert np.array_equal(ref.attr_global(attr, f), g)
        assert ref.attr_frame_count(attr) == f + 1
        last = p_ref
    slam.close()
