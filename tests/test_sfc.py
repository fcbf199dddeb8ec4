"""Safe Flight Corridor construction: the oracle's restatement of expandSFC / isObstacleInSFC against
first principles, then the device kernels (on the CPU emulator here, on the GPU in test_gpu_parity.py) against the oracle,
bit for bit.  PARITY UNPINNED against the reference: octomap / dynamicEDT3D are absent (see oracle/sfc_oracle.c)."""
import os

import numpy as np
import pytest

import emul
from lsc_dr_planner_b200 import missions as MS
from lsc_dr_planner_b200 import workloads as W
from oracle import oracle as orc

WORLD_DIR = "/root/reference/world"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def worlds():
    """(name, boxes): the committed copies of two reference worlds (tests/golden/worlds.npz, made by make_sfc_golden.py)"""
    g = np.load(os.path.join(GOLDEN, "worlds.npz"))
    return [(k, g[k]) for k in ("forest1", "maze1_dense")]


def unpack(c):
    return np.stack([c & 1023, (c >> 10) & 1023, (c >> 20) & 1023], -1)


def test_occupancy_follows_update_octree_from_csv():
    """MapManager::updateOctreeFromCSV (map_manager.cpp:262-305): cells round((c -+ s/2) / res) of every box"""
    boxes = np.array([[2.0, 2.0, 1.25, 0.5, 0.5, 2.5], [0.33, 3.01, 0.5, 0.2, 0.4, 1.0]])
    m = orc.Map(boxes, (0, 0, 0), (4, 4, 2.5))
    occ = m.occupancy()
    assert m.n == (41, 41, 26) and m.key0 == (0, 0, 0) and m.maxd2 == 100
    want = np.zeros_like(occ)
    want[18:23, 18:23, 0:25] = 1                       # round(17.5) = 18 .. round(22.5) = 23 (half away from zero)
    want[2:4, 28:32, 0:10] = 1                         # x: round(2.3) .. round(4.3); y: round(28.1) .. round(32.1)
    assert np.array_equal(occ, want)


def test_nearest_occupied_cell_is_euclidean_with_lowest_index_ties():
    boxes = np.array([[2.05, 2.05, 0.05, 0.1, 0.1, 0.1], [2.65, 2.05, 0.05, 0.1, 0.1, 0.1]])    # cells (20,20,0) and (26,20,0)
    m = orc.Map(boxes, (0, 0, 0), (4, 4, 1.0))
    cl = m.closest()
    assert tuple(cl[20, 20, 0]) == (20, 20, 0) and tuple(cl[22, 20, 0]) == (20, 20, 0) and tuple(cl[24, 20, 0]) == (26, 20, 0)
    assert tuple(cl[23, 20, 0]) == (20, 20, 0)         # equidistant: the lower x index
    assert tuple(cl[20, 29, 0]) == (20, 20, 0) and tuple(cl[20, 30, 0]) == (-1, -1, -1)       # valid strictly below 10 cells
    assert tuple(cl[20, 26, 8]) == (-1, -1, -1) and tuple(cl[20, 26, 7]) == (20, 20, 0)       # 36 + 64 = 100, 36 + 49 = 85


def _no_obstacle_within_margin(occ, box, res, margin):
    """brute force over every occupied cell: no grid point of the box lies within `margin` (L-inf) of an occupied cube"""
    idx = np.argwhere(occ)
    lo = idx * res; hi = (idx + 1) * res
    n = [int(np.floor((box[3 + k] - box[k] + 1e-5) / res)) + 1 for k in range(3)]
    g = np.stack(np.meshgrid(*[box[k] + np.arange(n[k]) * res for k in range(3)], indexing="ij"), -1).reshape(-1, 3)
    for pt in g:
        d = np.maximum(np.maximum(lo - pt, pt - hi), 0).max(axis=1)
        if d.min() < margin - 1e-6:
            return False
    return True


def test_expand_sfc_grows_to_a_maximal_obstacle_free_box():
    """expandSFC (:820-881) on a one-pillar world away from the origin: the grown box (before the margin compensation)
    keeps every grid point at least `margin` from the pillar, cannot grow by another cell on any face, and stays in the
    world; in an empty world it is the world itself"""
    res, margin = 0.1, 0.15
    empty = orc.Map(np.zeros((0, 6)), (2, 2, 0), (6, 6, 2))
    ok, box = empty.expand_sfc((3, 3, 1, 3, 3, 1), margin)
    assert ok and np.allclose(box, (2, 2, 0, 6, 6, 2), atol=2e-6)
    pillar = np.array([[4.0, 4.0, 1.0, 0.5, 0.5, 2.0]])
    m = orc.Map(pillar, (2, 2, 0), (6, 6, 2))
    ok, box = m.expand_sfc((3, 3, 1, 3.1, 3.1, 1.1), margin)
    assert ok
    delta = margin - int(margin / res) * res
    raw = box.astype(np.float64).copy()
    for k in range(3):                                  # undo the margin compensation (:868-877)
        if box[k] > (2, 2, 0)[k] + 1e-5:
            raw[k] += delta
        if box[3 + k] < (6, 6, 2)[k] - 1e-5:
            raw[3 + k] -= delta
    occ = m.occupancy()
    shifted = raw.copy(); shifted[:3] -= (2, 2, 0); shifted[3:] -= (2, 2, 0)
    assert _no_obstacle_within_margin(occ, shifted, res, margin)
    assert raw[0] <= 3 + 1e-6 and raw[3] >= 3.1 - 1e-6 and (raw[:3] >= np.array([2, 2, 0]) - 1e-5).all() and (raw[3:] <= np.array([6, 6, 2]) + 1e-5).all()
    for k in range(3):                                  # maximal: one more cell on any face meets the pillar or leaves the world
        for side in (0, 3):
            slab = raw.copy()
            if side == 0:
                slab[3 + k] = raw[k]; slab[k] = raw[k] - res
            else:
                slab[k] = raw[3 + k]; slab[3 + k] = raw[3 + k] + res
            outside = slab[k] < (2, 2, 0)[k] - 1e-5 or slab[3 + k] > (6, 6, 2)[k] + 1e-5
            assert outside or m.is_obstacle_in_sfc(slab, margin), (k, side, slab)
    assert not m.is_obstacle_in_sfc(raw, margin) and m.is_obstacle_in_sfc((3.5, 3.5, 0.5, 4.0, 4.0, 1.0), margin)


def test_phantom_cell_at_the_origin_when_no_obstacle_is_within_maxdist():
    """getDistanceAndClosestObstacle leaves the caller's default point3d (0,0,0) untouched when nothing is within
    maxdist, and isObstacleInSFC (:795-803) then measures against a cell there: restated as is"""
    m = orc.Map(np.zeros((0, 6)), (-2, -2, 0), (2, 2, 2))
    assert m.is_obstacle_in_sfc((-0.1, -0.1, 0.0, 0.1, 0.1, 0.1), 0.15)
    assert not m.is_obstacle_in_sfc((-0.1, -0.1, 0.3, 0.1, 0.1, 0.5), 0.15) and not m.is_obstacle_in_sfc((1.0, 1.0, 0.0, 1.2, 1.2, 0.1), 0.15)


def test_axis_order_prefers_the_goal_direction():
    """setAxisCand (:1134-1170): the face towards the largest goal offset grows first, its opposite last"""
    m = orc.Map(np.zeros((0, 6)), (2, 2, 0), (6, 6, 2))
    ok1, b1 = m.expand_sfc((3, 3, 1, 3.1, 3.1, 1.1), 0.15, goal=(5.5, 3.2, 1.0))
    ok2, b2 = m.expand_sfc((3, 3, 1, 3.1, 3.1, 1.1), 0.15)
    assert ok1 and ok2 and np.allclose(b1, b2, atol=2e-6)       # (an empty world ends at the world box either way)


def _agents_in_free_space(m, rng, n, world_min, world_max, radius):
    pts = []
    while len(pts) < n:
        p = rng.uniform(np.array(world_min) + 0.4, np.array(world_max) - 0.4).astype(np.float32)
        ok, _ = m.sfc_initialize(p, radius)
        if ok:
            pts.append(p)
    return np.array(pts, np.float32)


@pytest.mark.parametrize("which", [0, 1])
def test_device_map_and_sfc_kernels_match_the_oracle_on_reference_worlds(which):
    """occupancy_kernel == the oracle's grid; edt_closest_kernel == the oracle's nearest cells on sampled CTAs; sfc_kernel
    (initialize, from-point, from-convex-hull incl. the shift and the reuse-previous fallbacks) == the oracle bit for bit"""
    name, boxes = worlds()[which]
    wmin, wmax = (-5.0, -5.0, 0.0), (5.0, 5.0, 2.5)
    cfg = W.PlannerConfig(M=5, dim=3, world_min=wmin, world_max=wmax)
    m = orc.Map(boxes, wmin, wmax)
    em = emul.EmulMap(cfg, boxes)
    assert em.n == m.n and np.array_equal(em.occupancy(), m.occupancy())
    cl = m.closest()
    packed = np.where(cl[..., 0] < 0, -1, cl[..., 0] | (cl[..., 1] << 10) | (cl[..., 2] << 20)).astype(np.int32)
    rng = np.random.default_rng(which)
    cells = rng.choice(packed.size, 6, replace=False)
    em.edt_cells(cells)
    view = em.closest_view().reshape(-1)
    for c in cells:
        blk = slice((c // 128) * 128, min((c // 128 + 1) * 128, packed.size))
        assert np.array_equal(view[blk], packed.reshape(-1)[blk]), c
    em.closest_view()[...] = packed                      # (the full scan is too slow on the fiber emulator)
    n, radius, M = 6, 0.15, cfg.M
    pos = _agents_in_free_space(m, rng, n, wmin, wmax, radius)
    pos[0] = (4.0, 0.0, 1.0) if which == 0 else pos[0]
    lim = np.tile(np.array([1, 1, 1, 2, 2, 2, radius, 1.0]), (n, 1))
    sfc = np.zeros((n, M, 6), np.float32)
    st = em.sfc(0, pos, pos, pos, lim, sfc)
    for a in range(n):
        ok, box = m.sfc_initialize(pos[a], radius)
        assert st[a] == int(ok) == 1 and all(np.array_equal(sfc[a, s], box) for s in range(M)), (a, sfc[a, 0], box)
    # an agent inside an obstacle: reported, corridors untouched
    bad = np.array([boxes[0, :3]], np.float32); bad[0, 2] = 1.0
    sfc_bad = np.full((1, M, 6), 7.0, np.float32)
    assert em.sfc(0, bad, bad, bad, lim[:1], sfc_bad)[0] == 0 and (sfc_bad == 7.0).all() and not m.sfc_initialize(bad[0], radius)[0]
    # two closed-loop style updates per mode
    for mode in (1, 2):
        cur = sfc.copy()
        for step in range(2):
            last = (pos + rng.uniform(-0.4, 0.4, pos.shape)).astype(np.float32)
            goal = (last + rng.uniform(-1.0, 1.0, pos.shape)).astype(np.float32)
            wp = (goal + rng.uniform(-0.3, 0.3, pos.shape)).astype(np.float32)
            for arr in (last, goal, wp):
                arr[:, 2] = np.clip(arr[:, 2], 0.3, 2.2)
            want = cur.copy(); want_st = np.zeros(n, np.int32)
            for a in range(n):
                prev = cur[a, M - 1].copy()
                if mode == 1:
                    s_, box = m.sfc_from_point(last[a], goal[a], prev, radius)
                else:
                    s_, box = m.sfc_from_convex_hull([last[a], goal[a]], wp[a], prev, radius)
                want[a, :M - 1] = cur[a, 1:]; want[a, M - 1] = box; want_st[a] = s_
            got_st = em.sfc(mode, last, goal, wp, lim, cur)
            assert np.array_equal(got_st, want_st), (mode, step, got_st, want_st)
            assert np.array_equal(cur, want), (mode, step)
