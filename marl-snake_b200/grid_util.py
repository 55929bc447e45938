"""Wall layouts.  The reference ships text maps (marlenv/assets/*.txt: '#' wall, '.' empty) and a loader,
`make_grid_from_txt(map_path, mapper)` (core/grid_util.py:23-33); its SnakeEnv.reset always starts from `make_grid`'s
walled box (snake_env.py:133).  Here a layout is a constructor argument: `SnakeBatch(..., wall_map=...)` /
`SnakeEnv(..., wall_map=...)` / `make_snake(..., wall_map=...)`, given as an H x W array (nonzero = wall) or as the
path of such a text map."""
import numpy as np

DEFAULT_MAPPER = {'#': 1, '.': 0}


def make_grid_from_txt(map_path, mapper=None):
    """Same contract as the reference loader: one list of mapped characters per line -> integer array.  A trailing
    newline at the end of the file is tolerated (the reference would produce a ragged array there)."""
    mapper = DEFAULT_MAPPER if mapper is None else mapper
    with open(map_path, 'r') as fp:
        lines = fp.read().split('\n')
    while lines and lines[-1] == '':
        lines.pop()
    widths = {len(line) for line in lines}
    if len(widths) != 1:
        raise ValueError(f'{map_path}: rows have different lengths {sorted(widths)}')
    return np.asarray([[mapper[ch] for ch in line] for line in lines])


def as_wall_plane(wall_map, height=None, width=None):
    """wall_map (array-like H x W, or the path of a text map) -> contiguous uint8 [H, W] with 1 = wall."""
    if isinstance(wall_map, (str, bytes)) or hasattr(wall_map, '__fspath__'):
        wall_map = make_grid_from_txt(wall_map)
    plane = np.ascontiguousarray(np.asarray(wall_map) != 0, dtype=np.uint8)
    if plane.ndim != 2:
        raise ValueError('wall_map must be a 2-D array (H x W, nonzero = wall) or the path of a text map')
    if height is not None and (plane.shape[0] != height or plane.shape[1] != width):
        raise ValueError(f'wall_map is {plane.shape[0]}x{plane.shape[1]} but height x width is {height}x{width}')
    return plane
