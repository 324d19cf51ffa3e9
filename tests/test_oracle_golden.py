"""Pins the CPU oracle (oracle/vet_oracle.py) against fixtures produced by the
LIVE reference (tests/golden/make_golden.py).  Bit-exact for lattices, decode
tables, tile indices, counts; <= 1e-12 relative for entropies and weights."""
import numpy as np
import pytest

from conftest import group_keys, load_golden
from oracle import vet_oracle as orc

TILE_COUNTS_ALL = [20, 50, 100, 200, 250, 500, 1000]


def test_lattices_bit_exact():
    g = load_golden("lattices")
    for key in g.files:
        n = int(key[1:])
        L = orc.lattice(n)
        assert L.shape == g[key].shape == (orc.lattice_size(n), 3)
        assert np.array_equal(L, g[key]), key


def test_lattice_known_answers():
    # SURVEY Appendix B
    L = orc.lattice(20)
    assert tuple(L[0]) == (0.129235, -0.276168, -0.952381)
    assert tuple(L[10]) == (1.0, 0.0, 0.0)
    assert tuple(L[20]) == (0.129235, 0.276168, 0.952381)
    assert [orc.lattice_size(n) for n in (1, 2, 3, 100, 250, 500, 1000)] == [1, 3, 3, 101, 251, 501, 1001]
    L200 = orc.lattice(200)
    assert tuple(L200[0]) == (0.032803, 0.094072, -0.995025) and tuple(L200[100]) == (1.0, 0.0, 0.0)
    # mirror pairs c[2N-k] = (x,-y,-z)
    assert np.array_equal(L200[::-1] * np.array([1, -1, -1]), L200)


def test_axis_tables_and_cell_vectors_bit_exact():
    g = load_golden("decode")
    for (W, H) in [(100, 200), (200, 400), (64, 32), (1920, 1080), (3840, 1920)]:
        lon, lat = orc.axis_tables(W, H)
        assert np.array_equal(lon, g[f"lon_{W}x{H}"]), (W, H)
        assert np.array_equal(lat, g[f"lat_{W}x{H}"]), (W, H)
    for (W, H) in [(100, 200), (200, 400)]:
        cv = orc.cell_vectors(W, H)
        ref = g[f"cellvec_{W}x{H}"]
        assert np.array_equal(cv, ref)
        assert np.array_equal(np.signbit(cv), np.signbit(ref))  # -0.0 preserved


def test_decode_known_answers():
    g = load_golden("decode")
    mus = g["ntp_in"]
    assert np.array_equal(orc.normalize_to_pixel(mus, 100), g["ntp_100"])
    assert np.array_equal(orc.normalize_to_pixel(mus, 200), g["ntp_200"])
    assert np.array_equal(orc.normalize_to_pixel(mus, 1920), g["ntp_1920"])
    assert np.array_equal(orc.normalize_to_pixel(mus.astype(np.float32).astype(np.float64), 100), g["ntp32_100"])
    assert orc.normalize_to_pixel(np.array([0.29]), 100)[0] == 28  # trunc, not round
    # SURVEY Appendix B decode answers
    cv = orc.cell_vectors(100, 200)
    lon, lat = orc.axis_tables(100, 200)
    for mu, mv, px, py, lo, la, vec in [
        (0.5, 0.5, 50, 100, 0.0, 0.0, (1.0, 0.0, 0.0)),
        (0.0, 0.5, 0, 100, 0.0, 0.0, (1.0, 0.0, 0.0)),
        (1.0, 0.5, 100, 100, 180.0, 0.0, (-1.0, 0.0, 0.0)),
        (1.0, 1.0, 100, 200, 180.0, 0.0, (-1.0, 0.0, 0.0)),
        (0.29, 0.57, 28, 113, -79.2, -11.7, (0.183488, -0.961878, -0.202787)),
        (0.999, 0.001, 99, 0, 176.4, 90.0, (-0.0, 0.0, 1.0)),
        (0.123456, 0.654321, 12, 130, -136.8, -27.0, (-0.649516, -0.609936, -0.45399)),
        (0.75, 0.25, 75, 50, 90.0, 45.0, (0.0, 0.707107, 0.707107)),
    ]:
        qx, qy, ok = orc.decode(np.array([mu]), np.array([mv]), 100, 200)
        assert (qx[0], qy[0], bool(ok[0])) == (px, py, True)
        assert (lon[px], lat[py]) == (lo, la)
        assert tuple(cv[py, px]) == vec
    with pytest.raises(orc.OracleValidationError):
        orc.decode(np.array([1.5]), np.array([0.5]), 100, 200)
    with pytest.raises(orc.OracleValidationError):
        orc.axis_tables(101, 200)
    _, _, ok = orc.decode(np.array([np.nan, 0.5]), np.array([0.5, np.nan]), 100, 200)
    assert not ok.any()


def test_nearest_tile_exhaustive_default_grid():
    """All 20,301 reachable cells x seven tile counts == reference find_nearest_tile."""
    g = load_golden("nearest")
    cv = orc.cell_vectors(100, 200).reshape(-1, 3)
    for n in TILE_COUNTS_ALL:
        got = orc.nearest_tile(cv, orc.lattice(n))
        assert np.array_equal(got, g[f"lut_n{n}"].astype(np.int32)), n
    # exact ties at (-1,0,0) resolve to the lower index of a mirror pair
    v = np.array([[-1.0, 0.0, 0.0]])
    assert [int(orc.nearest_tile(v, orc.lattice(n))[0]) for n in (20, 50, 200)] == [6, 21, 83]


def test_nearest_tile_other_grid_and_arbitrary_vectors():
    g = load_golden("nearest")
    cv = orc.cell_vectors(200, 400).reshape(-1, 3)[g["sel_200x400"]]
    for n in (20, 200):
        assert np.array_equal(orc.nearest_tile(cv, orc.lattice(n)), g[f"lut200x400_n{n}"].astype(np.int32))
    for n in (20, 200, 1000):
        assert np.array_equal(orc.nearest_tile(g["arb_vecs"], orc.lattice(n)), g[f"arb_n{n}"].astype(np.int32))


def test_tile_weights_match_reference():
    g = load_golden("weights")
    cv = orc.cell_vectors(100, 200).reshape(-1, 3)[g["sel"]]
    for key in g.files:
        if not key.startswith("w_"):
            continue
        _, n, fov, pf = key.split("_")
        n, fov, pf = int(n[1:]), float(fov[3:]), float(pf[2:])
        w = orc.tile_weights(cv, orc.lattice(n), fov, True, pf)
        ref = g[key]
        assert np.array_equal(w > 0, ref > 0), key          # same support (d < fov/2)
        # weights just inside the threshold suffer cancellation in (max_d - d): absolute floor
        np.testing.assert_allclose(w, ref, rtol=1e-12, atol=1e-15, err_msg=key)


@pytest.mark.parametrize("case", ["c_small_w120", "c_small_unw", "c_w90_t200", "c_missing", "c_iid_unw", "c_iid_w",
                                  "c_oneuser", "c_t1000"])
def test_frame_fixtures(case):
    """compute_spatial_entropy / compute_transition_entropy of the reference on seeded
    packed tensors (decode chain included) vs the vectorised oracle."""
    c = group_keys(load_golden("frames"))[case]
    packed, tcs = c["packed"], [int(v) for v in c["tile_counts"]]
    fov, use_w, pf = float(c["cfg"][0]), bool(c["cfg"][1]), float(c["cfg"][2])
    sp = orc.spatial_analyzer(packed, 100, 200, tcs, fov, use_w, pf)
    assert np.array_equal(sp["assign0"], c["sp_assign0"])
    np.testing.assert_allclose(sp["per_k"], c["sp_per_k"], rtol=1e-12, atol=0, equal_nan=True)
    np.testing.assert_allclose(sp["entropy"], c["sp_entropy"], rtol=1e-12, atol=0, equal_nan=True)
    if use_w:
        np.testing.assert_allclose(sp["hist0"], c["sp_hist0"], rtol=1e-12, atol=0)
    else:
        assert np.array_equal(sp["hist0"], c["sp_hist0"])  # counts: exact
    tr = orc.transition_analyzer(packed, 100, 200, tcs)
    assert np.array_equal(tr["pairs0"], c["tr_pairs0"])
    assert np.array_equal(tr["prev_count0"], c["tr_prev_count0"])
    np.testing.assert_allclose(tr["per_k"], c["tr_per_k"], rtol=1e-12, atol=1e-15, equal_nan=True)
    np.testing.assert_allclose(tr["entropy"], c["tr_entropy"], rtol=1e-12, atol=1e-15, equal_nan=True)


def test_literal_layer_equals_vectorised_layer():
    """The scalar (reference-shaped) oracle layer and the vectorised layer agree."""
    c = group_keys(load_golden("frames"))["c_small_w120"]
    packed = c["packed"][:2, :10]
    vecs, ok = orc.decode_vectors(packed[..., 1], packed[..., 2], 100, 200)
    centres = orc.lattice(20)
    for f in range(2):
        d = {f"u{u}": tuple(vecs[f, u]) for u in range(10)}
        e, wts, asg = orc.compute_spatial_entropy_literal(d, centres, 120.0, True, 2.0)
        e2, hist, assign = orc.spatial_entropy(vecs[f], centres, 120.0, True, 2.0)
        assert list(asg.values()) == assign.tolist()
        np.testing.assert_allclose(e, e2, rtol=1e-12)
        for t, w in wts.items():
            np.testing.assert_allclose(hist[t], w, rtol=1e-12)
    d0 = {f"u{u}": tuple(vecs[0, u]) for u in range(10)}
    d1 = {f"u{u}": tuple(vecs[1, u]) for u in range(10)}
    e, m, pairs = orc.compute_transition_entropy_literal(d0, d1, centres)
    p = np.array([v[0] for v in pairs.values()]); cc = np.array([v[1] for v in pairs.values()])
    e2, m2 = orc.transition_entropy(p, cc, 21)
    np.testing.assert_allclose(e, e2, rtol=1e-12, atol=1e-15)
    assert {k: int(v) for k, v in m.items()} == {int(t): int(m2[t]) for t in np.flatnonzero(m2)}


def test_transition_quirk_fixtures():
    """Order-dependent bookkeeping (EU:278-318) on adversarial index patterns."""
    cases = group_keys(load_golden("transition_quirks"))
    assert len(cases) >= 40
    for name, c in cases.items():
        T = int(c["T"])
        e, _ = orc.transition_entropy(c["p"], c["c"], T)
        ref = float(c["e"])
        if np.isnan(ref):
            assert np.isnan(e), name
        else:
            np.testing.assert_allclose(e, ref, rtol=1e-12, atol=1e-15, err_msg=name)
    np.testing.assert_allclose(float(cases["appB"]["e"]), 0.41841441847669475, rtol=1e-14)


def test_transition_known_answers_appendix_b():
    prior = [(0.5, 0.5), (0.29, 0.57), (0.123456, 0.654321), (0.75, 0.25), (1.0, 0.5), (0.52, 0.48)]
    cur = [(0.51, 0.5), (0.30, 0.57), (0.123456, 0.654321), (0.70, 0.30), (0.99, 0.5), (0.52, 0.48)]
    packed = np.zeros((2, 6, 3))
    packed[0, :, 1:] = prior
    packed[1, :, 1:] = cur
    sp = orc.spatial_analyzer(packed[:1], 100, 200, [20], 120.0, True, 2.0)
    assert sp["assign0"][0].tolist() == [10, 8, 3, 17, 6, 10]
    np.testing.assert_allclose(sp["entropy"][0], 0.7617335881916677, rtol=1e-12)
    np.testing.assert_allclose(sp["hist0"][0].sum(), 6.001401199189895, rtol=1e-12)
    assert int((sp["hist0"][0] > 0).sum()) == 19
    sp200 = orc.spatial_analyzer(packed[:1], 100, 200, [200], 90.0, True, 2.0)
    assert sp200["assign0"][0].tolist() == [100, 77, 46, 170, 83, 113]
    np.testing.assert_allclose(sp200["entropy"][0], 0.8084991424139326, rtol=1e-12)
    un = orc.spatial_analyzer(packed[:1], 100, 200, [20, 200], 120.0, False, 2.0)
    np.testing.assert_allclose(un["per_k"][:, 0], [0.8710490642551527, 1.0], rtol=1e-12)
    tr = orc.transition_analyzer(packed, 100, 200, [20, 50, 200])
    np.testing.assert_allclose(tr["per_k"][:, 0], [0.1289509357448472, 0.1289509357448472, 0.0], rtol=1e-12, atol=1e-15)
    assert tr["pairs0"][0].tolist() == [[10, 10], [8, 8], [3, 3], [17, 12], [6, 14], [10, 10]]


def test_textbook_mode_differs_from_literal():
    p = np.full(12, 100); c = np.array([100, 100, 92, 92] * 3)
    lit, _ = orc.transition_entropy(p, c, 201)
    tb = orc.transition_entropy_textbook(p, c, 201)
    np.testing.assert_allclose(tb, 1.0 / np.log2(12), rtol=1e-12)
    assert abs(lit - tb) > 0.1


def test_edge_cases():
    centres = orc.lattice(20)
    with pytest.raises(orc.OracleValidationError):
        orc.compute_spatial_entropy_literal({}, centres)
    with pytest.raises(ZeroDivisionError):
        orc.transition_entropy(np.array([], dtype=int), np.array([], dtype=int), 21)
    e, _ = orc.transition_entropy(np.array([3]), np.array([4]), 21)
    assert np.isnan(e)  # one common user -> 0/0
    e1, _, _ = orc.spatial_entropy(np.array([[1.0, 0, 0]]), centres, 120.0, False, 2.0)
    assert np.isnan(e1)  # unweighted, one user
    e2, _, _ = orc.spatial_entropy(np.array([[1.0, 0, 0]]), orc.lattice(1), 120.0, True, 2.0)
    assert np.isnan(e2)  # T == 1
    packed = np.zeros((2, 3, 3)); packed[..., 1:] = 0.5; packed[1, :, 1] = np.nan
    with pytest.raises(orc.OracleValidationError):
        orc.spatial_analyzer(packed, 100, 200, [20])


# ---------------------------------------------------------------------------
# latitude/longitude grid tiling (NaiveSpatialEntropyAnalyzer) vs the live-reference fixtures
# ---------------------------------------------------------------------------
def _naive_groups():
    """{'cells/30x30': {...}, 'frames/n_30_w': {...}, 'analyzer/30x30_1': {...}} from 'a/b/field' keys."""
    npz = load_golden("naive")
    out = {}
    for k in npz.files:
        case, _, field = k.rpartition("/")
        out.setdefault(case, {})[field] = npz[k]
    return out


def test_naive_tile_index_every_cell():
    g = _naive_groups()
    lon_t, lat_t = orc.axis_tables(100, 200)
    n = 0
    for name, c in g.items():
        if not name.startswith("cells/"):
            continue
        tw, th = (int(x) for x in name.split("/")[1].split("x"))
        li, la = orc.naive_tile_index(lon_t[None, :], lat_t[:, None], tw, th)
        assert np.array_equal(np.broadcast_to(li, c["lon_idx"].shape), c["lon_idx"]), name
        assert np.array_equal(np.broadcast_to(la, c["lat_idx"].shape), c["lat_idx"]), name
        n += 1
    assert n == 6


def _naive_frames():
    return {k.split("/", 1)[1]: v for k, v in _naive_groups().items() if k.startswith("frames/")}


@pytest.mark.parametrize("case", ["n_30_w", "n_30_u", "n_45x90_u", "n_10x20_w", "n_360_u", "n_3_u", "n_120x60_one"])
def test_naive_frame_fixtures(case):
    c = _naive_frames()[case]
    tw, th = (int(x) for x in c["tile"])
    use_w = bool(c["use_w"])
    res = orc.naive_analyzer(c["packed"], 100, 200, tw, th, use_w)
    np.testing.assert_allclose(res["entropy"], c["entropy"], rtol=1e-12, atol=0, equal_nan=True)
    assert np.array_equal(res["lon_idx"], c["keys"][..., 0]) and np.array_equal(res["lat_idx"], c["keys"][..., 1])
    assert np.array_equal((res["hist0"] > 0).sum(1), c["nkeys"])
    # literal layer on the first frame
    px, py, ok = orc.decode(c["packed"][0, :, 1], c["packed"][0, :, 2], 100, 200)
    lon_t, lat_t = orc.axis_tables(100, 200)
    pts = {f"u{u:05d}": ((float(lon_t[px[u]]), float(lat_t[py[u]])) if ok[u] else None) for u in range(len(ok))}
    e, wts, asg = orc.compute_naive_spatial_entropy_literal(pts, th, tw, use_w)
    np.testing.assert_allclose(e, c["entropy"][0], rtol=1e-13, atol=0, equal_nan=True)
    assert len(wts) == c["nkeys"][0] and sum(wts.values()) == ok.sum()


def test_naive_validation():
    with pytest.raises(orc.OracleValidationError):
        orc.compute_naive_spatial_entropy_literal({}, 30, 30, True)
    with pytest.raises(orc.OracleValidationError):
        orc.compute_naive_spatial_entropy_literal({"a": (0.0, 0.0)}, 7, 30, True)
    with pytest.raises(orc.OracleValidationError):
        orc.compute_naive_spatial_entropy_literal({"a": (0.0, 0.0)}, 30, 7, True)
