"""Host-side visualisation of ONE environment of a batch: the grid is exported from the device
and drawn with NumPy.  Stands in for SnakeEnv.render_fancy / render('rgb_array') (snake_env.py:165-296,
grid_util.py:164-185) so RenderGUI and the eval / battle modes of the callers keep working; it is not on
the step path and is not pixel-identical to the reference's PIL drawing."""
import numpy as np

COLOR_BG = (40, 44, 52)
COLOR_WALL = (80, 80, 80)
COLOR_FRUIT = (230, 70, 70)
SNAKE_COLORS = [(80, 200, 120), (80, 160, 240), (200, 100, 240), (240, 200, 80)]


def rgb_from_grid(grid):
    """One pixel per cell (render('rgb_array'))."""
    grid = np.asarray(grid)
    out = np.zeros((*grid.shape, 3), dtype=np.uint8)
    kind, owner = grid % 10, grid // 10
    out[kind == 1] = (32, 32, 32)
    out[kind == 2] = (223, 7, 22)
    for k in (3, 4, 5):
        for i in np.unique(owner[kind == k]):
            base = np.array(SNAKE_COLORS[int(i) % len(SNAKE_COLORS)], dtype=np.float64)
            shade = base * (0.7 ** (int(i) // len(SNAKE_COLORS)))
            if k == 3:
                shade = np.minimum(255, shade * 1.3)
            out[(kind == k) & (owner == i)] = shade.astype(np.uint8)
    return out


def image_from_grid(grid, max_size=300):
    """PIL image of the grid, each cell scaled so that the longer side is about max_size pixels (one frame of
    render('gif'), grid_util.py:177-185)."""
    from PIL import Image
    grid = np.asarray(grid)
    scale = max(max_size // max(grid.shape), 1)
    rgb = np.repeat(np.repeat(rgb_from_grid(grid), scale, axis=0), scale, axis=1)
    return Image.fromarray(rgb, 'RGB')


def render_fancy(grid, heads=None, dirs=None, cell_size=40):
    """Upscaled frame: walls, round fruit, snake bodies, round heads with two eyes along the heading."""
    grid = np.asarray(grid)
    H, W = grid.shape
    cs = int(cell_size)
    img = np.empty((H * cs, W * cs, 3), dtype=np.uint8)
    img[:] = COLOR_BG
    yy, xx = np.mgrid[0:cs, 0:cs]
    centre = (cs - 1) / 2.0
    disc = (yy - centre) ** 2 + (xx - centre) ** 2 <= (cs * 0.5) ** 2
    small = (yy - centre) ** 2 + (xx - centre) ** 2 <= (cs * 0.3) ** 2
    kind, owner = grid % 10, grid // 10
    for r in range(H):
        for c in range(W):
            k = kind[r, c]
            if k == 0:
                continue
            tile = img[r * cs:(r + 1) * cs, c * cs:(c + 1) * cs]
            if k == 1:
                tile[:] = COLOR_WALL
            elif k == 2:
                tile[small] = COLOR_FRUIT
            else:
                color = SNAKE_COLORS[int(owner[r, c]) % len(SNAKE_COLORS)]
                if k == 3:
                    tile[disc] = color
                else:
                    tile[:] = color
    if heads is not None and dirs is not None:
        delta = ((-1, 0), (0, 1), (1, 0), (0, -1))
        eye = max(1, int(cs * 0.1))
        for h, d in zip(heads, dirs):
            if h < 0:
                continue
            r, c = divmod(int(h), W)
            dy, dx = delta[int(d)]
            cy, cx = r * cs + cs / 2.0, c * cs + cs / 2.0
            for side in (-1, 1):
                ey = int(cy + dy * cs * 0.3 + side * dx * cs * 0.15)
                ex = int(cx + dx * cs * 0.3 - side * dy * cs * 0.15)
                img[max(ey - eye, 0):ey + eye, max(ex - eye, 0):ex + eye] = (255, 255, 255)
    return img
