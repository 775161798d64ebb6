"""Oracle self-consistency on synthetic maps: documented edge semantics of the reference."""
import numpy as np

import synth


def test_hole_semantics(oracle):
    rows, cols = 64, 48
    z = synth.fbm(rows, cols, 0.02, 3, 0.15)
    z[20:26, 10:17] = np.nan
    z[22, 13] = np.float32(0.05)  # isolated valid cell inside the hole: nPoints == 1 in the roughness window
    z[40, 30] = np.inf
    g = oracle.Geometry.make(rows, cols, 0.02)
    o = oracle.chain(g, oracle.ChainParams.yaml_defaults(0), z, with_normals=True)
    hole = ~np.isfinite(z)
    for k in ("slope", "roughness", "traversability", "nx"):
        assert np.isnan(o[k][hole]).all(), k          # no normal at invalid cells
        assert np.isfinite(o[k][~hole]).all(), k
    assert np.isfinite(o["step"][21, 11])             # the step layer is also written at invalid cells (StepFilter.cpp:147)
    assert o["roughness"][22, 13] == 0.0              # 0/0 -> NaN -> comparison false -> 0.0 (RoughnessFilter.cpp:117-124)
    assert o["nz"][22, 13] == 1.0 and o["slope"][22, 13] == 1.0


def test_flat_and_tilted_planes(oracle):
    rows, cols = 40, 40
    g = oracle.Geometry.make(rows, cols, 0.02)
    p = oracle.ChainParams.yaml_defaults(0)
    flat = np.full((rows, cols), 0.25, np.float32)
    o = oracle.chain(g, p, flat)
    for k in ("slope", "step", "roughness", "traversability"):
        assert (o[k] == 1.0).all(), k
    # a plane with noise: slope equals the analytic inclination
    X = synth.cell_positions(rows, 0.02)[:, None]
    rng = np.random.default_rng(0)
    z = (0.3 * X + 1e-4 * rng.standard_normal((rows, cols))).astype(np.float32)
    o = oracle.chain(g, p, z, with_normals=True)
    th = np.arccos(o["nz"][5:-5, 5:-5].astype(np.float64))
    assert abs(np.median(th) - np.arctan(0.3)) < 2e-3
    assert (o["nx"][5:-5, 5:-5] < 0).all()           # normal leans against +x for a surface rising with x


def test_on_circle_membership_varies_with_position(oracle):
    # r = 0.04 at 0.02 m: offsets (+-2,0),(0,+-2) sit exactly on the circle; CircleIterator decides them on
    # absolute double positions, so the answer depends on the row
    g = oracle.Geometry.make(2048, 8, 0.02)
    counts = {len(oracle.circle_cells(g, i, 4, 0.04)[0]) for i in range(4, 2040, 7)}
    assert len(counts) > 1 and counts <= {9, 10, 11, 12, 13}
    g3 = oracle.Geometry.make(64, 64, 0.03)
    assert len(oracle.circle_cells(g3, 30, 30, 0.05)[0]) == 9
    assert len(oracle.circle_cells(g3, 0, 0, 0.05)[0]) == 4       # clipped at the map corner


def test_check_for_slope_threshold(oracle):
    # floor(2*(3*res)*(0.3/3)/res^2) = 29 at 0.02 m (the double expression is 29.999...), 20 at 0.03 m
    import math
    assert math.floor(2 * (3 * 0.02) * (0.3 / 3.0) / 0.02 ** 2) == 29
    assert math.floor(2 * (3 * 0.03) * (0.3 / 3.0) / 0.03 ** 2) == 20


def test_check_for_roughness_blocks_rough_patches(oracle):
    """checkForRoughness (TraversabilityMap.cpp:895-921): a cell of zero roughness-traversability is blocked when more than
    floor(1.5 * 3 res * (max_gap_width / 3) / res^2) = 22 cells of its 3 res circle (29 cells) are zero too."""
    rows, cols = 40, 40
    g = oracle.Geometry.make(rows, cols, 0.02)
    one = np.ones((rows, cols), np.float32, order="F")
    z = np.zeros((rows, cols), np.float32, order="F")
    rough = one.copy()
    rough[10:30, 10:30] = 0.0                       # a 20 x 20 patch of zero roughness traversability
    fp = oracle.FootprintParams.yaml_defaults()
    fp.radius, fp.offset = 0.0, 0.0                 # the sweep degenerates to the centre cell: blocked -> 0, else its traversability
    base, _, _ = oracle.footprint(g, fp, one, one, one, z)
    assert (base == 1.0).all()
    fp.verify_roughness = 1
    out, _, _, rfp = oracle.footprint(g, fp, one, one, one, z, roughness=rough)
    assert out[20, 20] == 0.0 and rfp[20, 20] == 0.0          # deep inside the patch: 29 zero cells > 22
    assert out[10, 10] == 1.0 and rfp[10, 10] == 1.0          # the patch corner sees only ~11 zero cells
    assert out[5, 5] == 1.0 and np.isnan(rfp[5, 5])           # not a zero-roughness cell: the check is skipped (:897)


def test_check_inclination_in_path_check(oracle):
    """checkInclination (TraversabilityMap.cpp:748-762) inside checkCircularFootprintPath (:359-363, :386-390): a zero of the
    robot_slope layer at a single pose, or on the grid line between two poses, makes the path unsafe; invalid cells are skipped."""
    rows, cols = 64, 64
    g = oracle.Geometry.make(rows, cols, 0.02)
    fp = np.asfortranarray(np.full((rows, cols), 0.8, dtype=np.float32))
    rs = np.asfortranarray(np.ones((rows, cols), dtype=np.float32))
    x = lambda i: 0.5 * rows * 0.02 - 0.01 - 0.02 * i   # cell centre of row index i (map centred at 0)
    y = lambda j: 0.5 * cols * 0.02 - 0.01 - 0.02 * j
    poses = np.array([[x(10), y(10)], [x(20), y(20)], [x(20), y(40)], [x(30), y(5)]])
    begin = [0, 1, 3, 4]                                # single pose, one segment along row 20, single pose
    safe, t = oracle.check_circular_paths(g, fp, 0.3, begin, poses, robot_slope=rs)
    assert safe.tolist() == [1, 1, 1] and np.allclose(t, 0.8)
    rs2 = rs.copy(); rs2[10, 10] = 0.0                  # at the first pose
    assert oracle.check_circular_paths(g, fp, 0.3, begin, poses, robot_slope=rs2)[0].tolist() == [0, 1, 1]
    rs3 = rs.copy(); rs3[20, 30] = 0.0                  # on the line of the segment
    assert oracle.check_circular_paths(g, fp, 0.3, begin, poses, robot_slope=rs3)[0].tolist() == [1, 0, 1]
    rs4 = rs.copy(); rs4[20, 30] = np.nan; rs4[21, 30] = 0.0   # an invalid cell on the line is skipped, a zero next to the line is not seen
    assert oracle.check_circular_paths(g, fp, 0.3, begin, poses, robot_slope=rs4)[0].tolist() == [1, 1, 1]
    assert oracle.check_circular_paths(g, fp, 0.3, begin, poses)[0].tolist() == [1, 1, 1]   # check off


def test_polygon_footprint_sweep_properties(oracle):
    """traversabilityFootprint(yaw) (TraversabilityMap.cpp:239-305, polygon isTraversable :592-645) on a constant layer: the mean of a
    constant is the constant wherever nothing is blocked, a blocked cell zeroes exactly the centres whose polygon covers it (a
    rectangle of the footprint's size for the unrotated polygon), yaw = 0 makes both layers equal, borders clip the polygon."""
    rows, cols, res = 96, 80, 0.03   # at 0.03 m checkForSlope can fail (more than 20 of the 29 window cells), at 0.02 m it cannot (29 of 29)
    g = oracle.Geometry.make(rows, cols, res)
    fp = oracle.FootprintParams.yaml_defaults()
    one = np.asfortranarray(np.ones((rows, cols), dtype=np.float32))
    trav = np.asfortranarray(np.full((rows, cols), 0.75, dtype=np.float32))
    z = np.asfortranarray(np.zeros((rows, cols), dtype=np.float32))
    poly = [[0.45, 0.30], [0.45, -0.30], [-0.45, -0.30], [-0.45, 0.30]]       # robot_footprint_parameter.yaml:3
    tx, trot = oracle.footprint_polygon(g, fp, poly, 0.7854, trav, one, one, z)
    assert np.all(tx == np.float32(0.75)) and np.all(trot == np.float32(0.75))
    tx0, trot0 = oracle.footprint_polygon(g, fp, poly, 0.0, trav, one, one, z)
    assert np.array_equal(tx0, trot0) and np.array_equal(tx0, tx)
    # blocked cells: slope == 0 in a patch large enough for checkForSlope (its inner cells see 29 zero cells within 3 cells)
    slope = one.copy(); slope[44:52, 36:44] = 0.0
    bx, brot = oracle.footprint_polygon(g, fp, poly, 0.7854, trav, slope, one, z)
    zero_x = np.argwhere(bx == 0.0)
    assert len(zero_x) > 0 and np.all((bx == 0.0) | (bx == np.float32(0.75)))
    # unrotated 0.9 m x 0.6 m rectangle: half extents 15 x 10 cells around every blocked cell
    assert zero_x[:, 0].min() >= 44 - 15 and zero_x[:, 0].max() <= 51 + 15
    assert zero_x[:, 1].min() >= 36 - 10 and zero_x[:, 1].max() <= 43 + 10
    assert (brot == 0.0).sum() > 0 and not np.array_equal(bx == 0.0, brot == 0.0)   # the rotated footprint covers other cells
