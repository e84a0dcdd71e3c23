"""Terrain generator.  The grid / curriculum / origin logic of ``Terrain`` is pinned against
the reference's own class (legged_gym/utils/terrain.py) run with OUR restated ``isaacgym.terrain_utils`` generators plugged
into the stub module: identical numpy seed -> bit-identical height field and env_origins.  The generators themselves are
third-party (Isaac Gym) and unpinned; they are checked through their geometric properties."""
import copy

import numpy as np
import pytest

from legged_games_gym_b200.envs import task_registry
from legged_games_gym_b200.utils import terrain as T
from oracle import ref_loader


def _cfg(**over):
    cfg = copy.deepcopy(task_registry.env_cfgs["anymal_c_rough"]).terrain
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("over", [dict(), dict(curriculum=False), dict(num_rows=3, num_cols=14, terrain_proportions=[0.1, 0.1, 0.2, 0.2, 0.1, 0.1, 0.1, 0.1]),
                                  dict(mesh_type="heightfield", curriculum=False, num_rows=4, num_cols=4)])
def test_terrain_class_matches_reference_class(over):
    ref_loader.load_reference()
    import sys
    stub = sys.modules["isaacgym.terrain_utils"]
    for name in ("SubTerrain", "random_uniform_terrain", "sloped_terrain", "pyramid_sloped_terrain", "discrete_obstacles_terrain",
                 "wave_terrain", "stairs_terrain", "pyramid_stairs_terrain", "stepping_stones_terrain"):
        setattr(stub, name, getattr(T, name))
    # the trimesh conversion is checked separately below; keep the reference's eager call cheap
    stub.convert_heightfield_to_trimesh = lambda *a, **k: (None, None)
    import legged_gym.utils.terrain as RT
    RT.terrain_utils = stub
    np.random.seed(11)
    ref = RT.Terrain(_cfg(**over), 64)
    np.random.seed(11)
    got = T.Terrain(_cfg(**over), 64)
    assert got.tot_rows == ref.tot_rows and got.tot_cols == ref.tot_cols and got.border == ref.border
    assert got.heightsamples.dtype == np.int16
    assert np.array_equal(got.heightsamples, ref.heightsamples)
    assert np.array_equal(got.env_origins, ref.env_origins)


def test_default_terrain_geometry():
    cfg = _cfg()
    np.random.seed(0)
    t = T.Terrain(cfg, 4096)
    assert t.heightsamples.shape == (1300, 2100) and t.env_origins.shape == (10, 20, 3)
    b = t.border
    assert not t.heightsamples[:b].any() and not t.heightsamples[-b:].any() and not t.heightsamples[:, :b].any()
    # origins sit at the cell centres; z is the platform height in metres
    assert np.allclose(t.env_origins[:, :, 0], ((np.arange(10) + 0.5) * 8.0)[:, None])
    assert np.allclose(t.env_origins[:, :, 1], ((np.arange(20) + 0.5) * 8.0)[None, :])
    # curriculum: difficulty grows with the row -> the stairs-up columns (band 3) get higher platforms row by row
    props = np.cumsum(cfg.terrain_proportions)
    col = next(j for j in range(20) if props[2] <= j / 20 + 0.001 < props[3])
    z = t.env_origins[:, col, 2]
    assert np.all(np.diff(z) > 0)
    # deterministic under the numpy seed
    np.random.seed(0)
    assert np.array_equal(T.Terrain(_cfg(), 4096).heightsamples, t.heightsamples)


def test_generators_shapes_and_platforms():
    def sub():
        return T.SubTerrain("terrain", width=80, length=80, vertical_scale=0.005, horizontal_scale=0.1)
    s = T.pyramid_stairs_terrain(sub(), step_width=0.31, step_height=0.1, platform_size=3.)
    assert s.height_field_raw[0, 0] == 0 and s.height_field_raw[40, 40] == s.height_field_raw.max() > 0
    assert set(np.unique(np.diff(np.unique(s.height_field_raw)))) == {20}             # equal risers of 0.1 m / 0.005
    p = T.pyramid_sloped_terrain(sub(), slope=0.2, platform_size=3.)
    assert p.height_field_raw[40, 40] == p.height_field_raw.max() and p.height_field_raw[0, 0] == 0
    n = T.pyramid_sloped_terrain(sub(), slope=-0.2, platform_size=3.)
    assert n.height_field_raw.min() < 0 and n.height_field_raw.max() == 0
    np.random.seed(1)
    r = T.random_uniform_terrain(sub(), min_height=-0.05, max_height=0.05, step=0.005, downsampled_scale=0.2)
    assert -10 <= r.height_field_raw.min() and r.height_field_raw.max() <= 10 and r.height_field_raw.std() > 1
    d = T.discrete_obstacles_terrain(sub(), 0.2, 1., 2., 20, platform_size=3.)
    assert not d.height_field_raw[25:55, 25:55].any() and set(np.unique(d.height_field_raw)) <= {-40, -20, 0, 20, 40}
    st = T.stepping_stones_terrain(sub(), stone_size=1.0, stone_distance=0.1, max_height=0., platform_size=4.)
    assert st.height_field_raw.min() == int(-10 / 0.005) and not st.height_field_raw[20:60, 20:60].any()
    g = sub(); T.gap_terrain(g, gap_size=0.5, platform_size=3.)
    assert g.height_field_raw.min() == -1000 and g.height_field_raw[40, 40] == 0
    q = sub(); T.pit_terrain(q, depth=0.5, platform_size=4.)
    assert q.height_field_raw.min() == -100 and q.height_field_raw[0, 0] == 0


def test_trimesh_conversion_small():
    hf = np.array([[0, 0, 0], [0, 40, 0], [0, 0, 0]], dtype=np.int16)
    v, tri = T.convert_heightfield_to_trimesh(hf, 0.1, 0.005, slope_threshold=0.75)
    assert v.shape == (9, 3) and tri.shape == (8, 3) and tri.dtype == np.uint32
    assert np.isclose(v[4, 2], 0.2) and tri.max() == 8
    # the steep centre spike pulls its neighbours' xy onto it (vertical walls)
    v2, _ = T.convert_heightfield_to_trimesh(hf, 0.1, 0.005, slope_threshold=None)
    assert not np.allclose(v[:, :2], v2[:, :2])
    cfg = _cfg(num_rows=1, num_cols=1, border_size=1.0)
    np.random.seed(2)
    t = T.Terrain(cfg, 8)
    assert t.vertices.shape == (t.tot_rows * t.tot_cols, 3) and t.triangles.shape == (2 * (t.tot_rows - 1) * (t.tot_cols - 1), 3)
