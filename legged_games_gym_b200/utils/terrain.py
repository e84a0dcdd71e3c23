"""Terrain generation (reference legged_gym/utils/terrain.py:38-187).

The reference builds its height field with ``isaacgym.terrain_utils``, a module of the closed Isaac Gym package that is
not under /root/reference.  The sub-terrain generators below restate its published algorithms (Isaac Gym Preview 3,
python/isaacgym/terrain_utils.py: SubTerrain, random_uniform_terrain, sloped_terrain, pyramid_sloped_terrain,
discrete_obstacles_terrain, wave_terrain, stairs_terrain, pyramid_stairs_terrain, stepping_stones_terrain,
convert_heightfield_to_trimesh); PARITY UNPINNED for those (no test or fixture of them exists in the reference).
``Terrain`` itself, ``gap_terrain`` and ``pit_terrain`` follow the reference file line by line in behaviour: same grid
arithmetic, same curriculum (difficulty = row / num_rows, type = col / num_cols + 0.001), same ``env_origins``.

Init-time host code (numpy); the per-step consumer is the height-scan kernel, which reads ``heightsamples`` (int16).
"""
import numpy as np


class SubTerrain:
    def __init__(self, terrain_name="terrain", width=256, length=256, vertical_scale=1.0, horizontal_scale=1.0):
        self.terrain_name = terrain_name
        self.vertical_scale = vertical_scale
        self.horizontal_scale = horizontal_scale
        self.width = width
        self.length = length
        self.height_field_raw = np.zeros((self.width, self.length), dtype=np.int16)


def _bilinear_resample(coarse, n_rows, n_cols):
    """linear interpolation of a regular grid onto n_rows x n_cols points spanning the same extent (what
    scipy.interpolate.interp2d(kind='linear') computed for terrain_utils; interp2d no longer exists in scipy)."""
    r = np.linspace(0.0, coarse.shape[0] - 1.0, n_rows)
    c = np.linspace(0.0, coarse.shape[1] - 1.0, n_cols)
    r0 = np.clip(np.floor(r).astype(int), 0, max(coarse.shape[0] - 2, 0))
    c0 = np.clip(np.floor(c).astype(int), 0, max(coarse.shape[1] - 2, 0))
    r1 = np.minimum(r0 + 1, coarse.shape[0] - 1)
    c1 = np.minimum(c0 + 1, coarse.shape[1] - 1)
    fr = (r - r0)[:, None]
    fc = (c - c0)[None, :]
    z = coarse.astype(np.float64)
    top = z[r0][:, c0] * (1 - fc) + z[r0][:, c1] * fc
    bot = z[r1][:, c0] * (1 - fc) + z[r1][:, c1] * fc
    return top * (1 - fr) + bot * fr


def random_uniform_terrain(terrain, min_height, max_height, step=1, downsampled_scale=None):
    if downsampled_scale is None:
        downsampled_scale = terrain.horizontal_scale
    min_height = int(min_height / terrain.vertical_scale)
    max_height = int(max_height / terrain.vertical_scale)
    step = int(step / terrain.vertical_scale)
    heights_range = np.arange(min_height, max_height + step, step)
    shape = (int(terrain.width * terrain.horizontal_scale / downsampled_scale),
             int(terrain.length * terrain.horizontal_scale / downsampled_scale))
    coarse = np.random.choice(heights_range, shape)
    z = np.rint(_bilinear_resample(coarse, terrain.width, terrain.length))
    terrain.height_field_raw += z.astype(np.int16)
    return terrain


def sloped_terrain(terrain, slope=1):
    xx = np.arange(0, terrain.width).reshape(terrain.width, 1)
    max_height = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * terrain.width)
    terrain.height_field_raw[:, np.arange(terrain.length)] += (max_height * xx / terrain.width).astype(terrain.height_field_raw.dtype)
    return terrain


def pyramid_sloped_terrain(terrain, slope=1, platform_size=1.):
    x = np.arange(0, terrain.width)
    y = np.arange(0, terrain.length)
    center_x = int(terrain.width / 2)
    center_y = int(terrain.length / 2)
    xx = ((center_x - np.abs(center_x - x)) / center_x).reshape(terrain.width, 1)
    yy = ((center_y - np.abs(center_y - y)) / center_y).reshape(1, terrain.length)
    max_height = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * (terrain.width / 2))
    terrain.height_field_raw += (max_height * xx * yy).astype(terrain.height_field_raw.dtype)
    platform_size = int(platform_size / terrain.horizontal_scale / 2)
    x1 = terrain.width // 2 - platform_size
    x2 = terrain.width // 2 + platform_size
    y1 = terrain.length // 2 - platform_size
    y2 = terrain.length // 2 + platform_size
    min_h = min(terrain.height_field_raw[x1, y1], 0)
    max_h = max(terrain.height_field_raw[x1, y1], 0)
    terrain.height_field_raw = np.clip(terrain.height_field_raw, min_h, max_h)
    return terrain


def discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size=1.):
    max_height = int(max_height / terrain.vertical_scale)
    min_size = int(min_size / terrain.horizontal_scale)
    max_size = int(max_size / terrain.horizontal_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    (i, j) = terrain.height_field_raw.shape
    height_range = [-max_height, -max_height // 2, max_height // 2, max_height]
    width_range = range(min_size, max_size, 4)
    length_range = range(min_size, max_size, 4)
    for _ in range(num_rects):
        width = np.random.choice(width_range)
        length = np.random.choice(length_range)
        start_i = np.random.choice(range(0, i - width, 4))
        start_j = np.random.choice(range(0, j - length, 4))
        terrain.height_field_raw[start_i:start_i + width, start_j:start_j + length] = np.random.choice(height_range)
    x1 = (terrain.width - platform_size) // 2
    x2 = (terrain.width + platform_size) // 2
    y1 = (terrain.length - platform_size) // 2
    y2 = (terrain.length + platform_size) // 2
    terrain.height_field_raw[x1:x2, y1:y2] = 0
    return terrain


def wave_terrain(terrain, num_waves=1, amplitude=1.):
    amplitude = int(0.5 * amplitude / terrain.vertical_scale)
    if num_waves > 0:
        div = terrain.length / (num_waves * np.pi * 2)
        xx = np.arange(0, terrain.width).reshape(terrain.width, 1)
        yy = np.arange(0, terrain.length).reshape(1, terrain.length)
        terrain.height_field_raw += (amplitude * np.cos(yy / div) + amplitude * np.sin(xx / div)).astype(terrain.height_field_raw.dtype)
    return terrain


def stairs_terrain(terrain, step_width, step_height):
    step_width = int(step_width / terrain.horizontal_scale)
    step_height = int(step_height / terrain.vertical_scale)
    num_steps = terrain.width // step_width
    height = step_height
    for i in range(num_steps):
        terrain.height_field_raw[i * step_width:(i + 1) * step_width, :] += height
        height += step_height
    return terrain


def pyramid_stairs_terrain(terrain, step_width, step_height, platform_size=1.):
    step_width = int(step_width / terrain.horizontal_scale)
    step_height = int(step_height / terrain.vertical_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    height = 0
    start_x, stop_x, start_y, stop_y = 0, terrain.width, 0, terrain.length
    while (stop_x - start_x) > platform_size and (stop_y - start_y) > platform_size:
        start_x += step_width
        stop_x -= step_width
        start_y += step_width
        stop_y -= step_width
        height += step_height
        terrain.height_field_raw[start_x:stop_x, start_y:stop_y] = height
    return terrain


def stepping_stones_terrain(terrain, stone_size, stone_distance, max_height, platform_size=1., depth=-10):
    stone_size = int(stone_size / terrain.horizontal_scale)
    stone_distance = int(stone_distance / terrain.horizontal_scale)
    max_height = int(max_height / terrain.vertical_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    height_range = np.arange(-max_height - 1, max_height, step=1)
    start_x, start_y = 0, 0
    terrain.height_field_raw[:, :] = int(depth / terrain.vertical_scale)
    if terrain.length >= terrain.width:
        while start_y < terrain.length:
            stop_y = min(terrain.length, start_y + stone_size)
            start_x = np.random.randint(0, stone_size)
            stop_x = max(0, start_x - stone_distance)                       # fill first hole
            terrain.height_field_raw[0:stop_x, start_y:stop_y] = np.random.choice(height_range)
            while start_x < terrain.width:                                  # fill row
                stop_x = min(terrain.width, start_x + stone_size)
                terrain.height_field_raw[start_x:stop_x, start_y:stop_y] = np.random.choice(height_range)
                start_x += stone_size + stone_distance
            start_y += stone_size + stone_distance
    else:
        while start_x < terrain.width:
            stop_x = min(terrain.width, start_x + stone_size)
            start_y = np.random.randint(0, stone_size)
            stop_y = max(0, start_y - stone_distance)
            terrain.height_field_raw[start_x:stop_x, 0:stop_y] = np.random.choice(height_range)
            while start_y < terrain.length:
                stop_y = min(terrain.length, start_y + stone_size)
                terrain.height_field_raw[start_x:stop_x, start_y:stop_y] = np.random.choice(height_range)
                start_y += stone_size + stone_distance
            start_x += stone_size + stone_distance
    x1 = (terrain.width - platform_size) // 2
    x2 = (terrain.width + platform_size) // 2
    y1 = (terrain.length - platform_size) // 2
    y2 = (terrain.length + platform_size) // 2
    terrain.height_field_raw[x1:x2, y1:y2] = 0
    return terrain


def convert_heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    """vertices [rows*cols, 3] float32, triangles [2(rows-1)(cols-1), 3] uint32; slopes steeper than the threshold are
    made vertical by moving the upper vertices onto the lower ones' xy."""
    hf = height_field_raw
    num_rows, num_cols = hf.shape
    y = np.linspace(0, (num_cols - 1) * horizontal_scale, num_cols)
    x = np.linspace(0, (num_rows - 1) * horizontal_scale, num_rows)
    yy, xx = np.meshgrid(y, x)
    if slope_threshold is not None:
        slope_threshold *= horizontal_scale / vertical_scale
        move_x = np.zeros((num_rows, num_cols))
        move_y = np.zeros((num_rows, num_cols))
        move_corners = np.zeros((num_rows, num_cols))
        h = hf.astype(np.int64)
        move_x[:num_rows - 1, :] += (h[1:num_rows, :] - h[:num_rows - 1, :] > slope_threshold)
        move_x[1:num_rows, :] -= (h[:num_rows - 1, :] - h[1:num_rows, :] > slope_threshold)
        move_y[:, :num_cols - 1] += (h[:, 1:num_cols] - h[:, :num_cols - 1] > slope_threshold)
        move_y[:, 1:num_cols] -= (h[:, :num_cols - 1] - h[:, 1:num_cols] > slope_threshold)
        move_corners[:num_rows - 1, :num_cols - 1] += (h[1:num_rows, 1:num_cols] - h[:num_rows - 1, :num_cols - 1] > slope_threshold)
        move_corners[1:num_rows, 1:num_cols] -= (h[:num_rows - 1, :num_cols - 1] - h[1:num_rows, 1:num_cols] > slope_threshold)
        xx += (move_x + move_corners * (move_x == 0)) * horizontal_scale
        yy += (move_y + move_corners * (move_y == 0)) * horizontal_scale
    vertices = np.zeros((num_rows * num_cols, 3), dtype=np.float32)
    vertices[:, 0] = xx.flatten()
    vertices[:, 1] = yy.flatten()
    vertices[:, 2] = hf.flatten() * vertical_scale
    triangles = -np.ones((2 * (num_rows - 1) * (num_cols - 1), 3), dtype=np.int64)
    for i in range(num_rows - 1):
        ind0 = np.arange(0, num_cols - 1) + i * num_cols
        ind1 = ind0 + 1
        ind2 = ind0 + num_cols
        ind3 = ind2 + 1
        start = 2 * i * (num_cols - 1)
        stop = start + 2 * (num_cols - 1)
        triangles[start:stop:2, 0] = ind0
        triangles[start:stop:2, 1] = ind3
        triangles[start:stop:2, 2] = ind1
        triangles[start + 1:stop:2, 0] = ind0
        triangles[start + 1:stop:2, 1] = ind2
        triangles[start + 1:stop:2, 2] = ind3
    return vertices, triangles.astype(np.uint32)


class Terrain:
    """Grid of num_rows x num_cols sub-terrains inside a flat border (reference utils/terrain.py:38-164).
    Attributes read by the env: cfg, env_length, env_width, border, tot_rows, tot_cols, heightsamples (int16
    [tot_rows, tot_cols]), env_origins [num_rows, num_cols, 3], and for 'trimesh' vertices / triangles."""

    def __init__(self, cfg, num_robots) -> None:
        self.cfg, self.num_robots, self.type = cfg, num_robots, cfg.mesh_type
        if self.type in ("none", "plane"):
            return
        hs = cfg.horizontal_scale
        self.env_length, self.env_width = cfg.terrain_length, cfg.terrain_width
        self.proportions = [np.sum(cfg.terrain_proportions[:k + 1]) for k in range(len(cfg.terrain_proportions))]
        cfg.num_sub_terrains = cfg.num_rows * cfg.num_cols
        self.width_per_env_pixels = int(self.env_width / hs)
        self.length_per_env_pixels = int(self.env_length / hs)
        self.border = int(cfg.border_size / hs)
        self.tot_rows = int(cfg.num_rows * self.length_per_env_pixels) + 2 * self.border
        self.tot_cols = int(cfg.num_cols * self.width_per_env_pixels) + 2 * self.border
        self.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        if cfg.curriculum:
            self.curiculum()
        elif cfg.selected:
            self.selected_terrain()
        else:
            self.randomized_terrain()
        self.heightsamples = self.height_field_raw
        self._mesh = None

    def _trimesh(self):
        # the reference converts eagerly (terrain.py:67-71) to hand the mesh to PhysX; nothing on the hot path reads it,
        # so the 2.7 M-vertex conversion runs on first access
        if self._mesh is None:
            if self.type != "trimesh":
                raise AttributeError("vertices / triangles exist for mesh_type 'trimesh' only")
            self._mesh = convert_heightfield_to_trimesh(self.height_field_raw, self.cfg.horizontal_scale,
                                                        self.cfg.vertical_scale, self.cfg.slope_treshold)
        return self._mesh

    @property
    def vertices(self):
        return self._trimesh()[0]

    @property
    def triangles(self):
        return self._trimesh()[1]

    def _blank(self):
        # the reference passes width_per_env_pixels for BOTH sides (terrain.py:112-116): sub-terrains are square
        return SubTerrain("terrain", width=self.width_per_env_pixels, length=self.width_per_env_pixels,
                          vertical_scale=self.cfg.vertical_scale, horizontal_scale=self.cfg.horizontal_scale)

    def randomized_terrain(self):
        """every cell: uniform type draw, difficulty from {0.5, 0.75, 0.9} (terrain.py:74-83; numpy global RNG)"""
        for k in range(self.cfg.num_sub_terrains):
            row, col = np.unravel_index(k, (self.cfg.num_rows, self.cfg.num_cols))
            choice = np.random.uniform(0, 1)
            difficulty = np.random.choice([0.5, 0.75, 0.9])
            self.add_terrain_to_map(self.make_terrain(choice, difficulty), row, col)

    def curiculum(self):
        """difficulty grows with the row, the type is fixed per column (terrain.py:85-93; reference spelling kept)"""
        for col in range(self.cfg.num_cols):
            for row in range(self.cfg.num_rows):
                self.add_terrain_to_map(self.make_terrain(col / self.cfg.num_cols + 0.001, row / self.cfg.num_rows), row, col)

    def selected_terrain(self):
        """one generator, named by cfg.terrain_kwargs['type'], for every cell (terrain.py:95-109)"""
        kwargs = dict(self.cfg.terrain_kwargs)
        name = kwargs.pop("type").split(".")[-1]
        table = {f.__name__: f for f in (random_uniform_terrain, sloped_terrain, pyramid_sloped_terrain,
                                         discrete_obstacles_terrain, wave_terrain, stairs_terrain,
                                         pyramid_stairs_terrain, stepping_stones_terrain, gap_terrain, pit_terrain)}
        for k in range(self.cfg.num_sub_terrains):
            row, col = np.unravel_index(k, (self.cfg.num_rows, self.cfg.num_cols))
            sub = self._blank()
            table[name](sub, **kwargs)
            self.add_terrain_to_map(sub, row, col)

    def make_terrain(self, choice, difficulty):
        """type = first cumulative proportion above `choice` (terrain.py:111-149): 0 smooth slope (downhill for the
        lower half of its band), 1 rough slope, 2 stairs down, 3 stairs up, 4 discrete obstacles, 5 stepping stones,
        6 gap, beyond: pit.  All sizes scale with `difficulty` exactly as in the reference."""
        sub = self._blank()
        band = next((k for k, edge in enumerate(self.proportions) if choice < edge), len(self.proportions))
        slope = difficulty * 0.4
        if band <= 1:
            if band == 0 and choice < self.proportions[0] / 2:
                slope = -slope
            pyramid_sloped_terrain(sub, slope=slope, platform_size=3.)
            if band == 1:
                random_uniform_terrain(sub, min_height=-0.05, max_height=0.05, step=0.005, downsampled_scale=0.2)
        elif band <= 3:
            rise = 0.05 + 0.18 * difficulty
            pyramid_stairs_terrain(sub, step_width=0.31, step_height=(-rise if band == 2 else rise), platform_size=3.)
        elif band == 4:
            discrete_obstacles_terrain(sub, 0.05 + difficulty * 0.2, 1., 2., 20, platform_size=3.)
        elif band == 5:
            stepping_stones_terrain(sub, stone_size=1.5 * (1.05 - difficulty), stone_distance=(0.05 if difficulty == 0 else 0.1),
                                    max_height=0., platform_size=4.)
        elif band == 6:
            gap_terrain(sub, gap_size=1. * difficulty, platform_size=3.)
        else:
            pit_terrain(sub, depth=1. * difficulty, platform_size=4.)
        return sub

    def add_terrain_to_map(self, terrain, row, col):
        """paste the cell and set its spawn origin: cell centre, z = highest sample within +-1 m of it (terrain.py:151-170)"""
        r0 = self.border + row * self.length_per_env_pixels
        c0 = self.border + col * self.width_per_env_pixels
        self.height_field_raw[r0:r0 + self.length_per_env_pixels, c0:c0 + self.width_per_env_pixels] = terrain.height_field_raw
        hs = terrain.horizontal_scale
        xs = slice(int((self.env_length / 2. - 1) / hs), int((self.env_length / 2. + 1) / hs))
        ys = slice(int((self.env_width / 2. - 1) / hs), int((self.env_width / 2. + 1) / hs))
        z = np.max(terrain.height_field_raw[xs, ys]) * terrain.vertical_scale
        self.env_origins[row, col] = [(row + 0.5) * self.env_length, (col + 0.5) * self.env_width, z]


def gap_terrain(terrain, gap_size, platform_size=1.):
    gap_size = int(gap_size / terrain.horizontal_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    center_x = terrain.length // 2
    center_y = terrain.width // 2
    x1 = (terrain.length - platform_size) // 2
    x2 = x1 + gap_size
    y1 = (terrain.width - platform_size) // 2
    y2 = y1 + gap_size
    terrain.height_field_raw[center_x - x2:center_x + x2, center_y - y2:center_y + y2] = -1000
    terrain.height_field_raw[center_x - x1:center_x + x1, center_y - y1:center_y + y1] = 0


def pit_terrain(terrain, depth, platform_size=1.):
    depth = int(depth / terrain.vertical_scale)
    platform_size = int(platform_size / terrain.horizontal_scale / 2)
    x1 = terrain.length // 2 - platform_size
    x2 = terrain.length // 2 + platform_size
    y1 = terrain.width // 2 - platform_size
    y2 = terrain.width // 2 + platform_size
    terrain.height_field_raw[x1:x2, y1:y2] = -depth
