"""Host-side visualisation of ONE environment of a batch: the state is exported from the device
(`snk_get_state`) and drawn here, pixel for pixel what the reference draws -- `SnakeEnv.render_fancy`
(snake_env.py:165-263), `render('rgb_array')` (core/grid_util.py:164-174) and the frames of `render('gif')`
(:177-185) -- so RenderGUI windows, videos and GIFs of the callers look the same.  Not on the step path.

tests/golden/render.npz holds frames of the unmodified reference; tests/test_render.py compares against them."""
import numpy as np

# snake_env.py:20-30
COLOR_BG = (40, 44, 52)
COLOR_WALL = (80, 80, 80)
COLOR_FRUIT = (230, 70, 70)
SNAKE_COLORS = [(80, 200, 120), (80, 160, 240), (200, 100, 240), (240, 200, 80)]

# core/snake.py:14-31 (CellColors): one pixel per cell; head = the body colour doubled and clipped
BODY_WHEEL = [(104, 255, 0), (255, 191, 0), (255, 0, 92), (0, 111, 255)]
HEAD_WHEEL = [tuple(min(255, int(v * 2.0)) for v in rgb) for rgb in BODY_WHEEL]
CELL_COLORS = {0: [(0, 0, 0)], 1: [(32, 32, 32)], 2: [(223, 7, 22)], 3: HEAD_WHEEL, 4: BODY_WHEEL, 5: BODY_WHEEL}
DELTA = ((-1, 0), (0, 1), (1, 0), (0, -1))          # direction codes 0 UP 1 RIGHT 2 DOWN 3 LEFT as (dy, dx)


def rgb_from_grid(grid):
    """render('rgb_array'): uint8 [H, W, 3], one pixel per cell.  Colour = wheel[owner % 4] of the cell type, dimmed
    by 0.7 per full turn of the wheel (owners 4..7 once, 8..11 twice, ...), truncated to uint8."""
    grid = np.asarray(grid).astype(np.int64)
    kind, owner = grid % 10, grid // 10
    out = np.zeros((*grid.shape, 3), dtype=np.uint8)
    for k, wheel in CELL_COLORS.items():
        table = np.array(wheel, dtype=np.float64)
        sel = kind == k
        if not sel.any():
            continue
        who = owner[sel]
        out[sel] = (table[who % len(wheel)] * (0.7 ** (who // len(wheel)))[:, None]).astype(np.uint8)
    return out


def image_from_grid(grid, max_size=300):
    """One frame of render('gif'): the rgb_array picture with every cell blown up to max(max_size // longer side, 1)
    pixels, as a PIL image."""
    from PIL import Image
    grid = np.asarray(grid)
    scale = max(max_size // max(grid.shape), 1)
    return Image.fromarray(np.repeat(np.repeat(rgb_from_grid(grid), scale, axis=0), scale, axis=1), 'RGB')


def fancy_ops(grid, cells, alive, dirs, cell_size):
    """The drawing as a list of primitives ('rect' | 'oval', box, colour) in paint order: walls and fruits in
    row-major order, then every live snake in index order -- one square per body cell (head included), a disc on the
    head cell, two white eyes with black pupils placed 0.3 cells ahead of and 0.15 cells beside the head centre.
    Boxes are PIL boxes: the far edge is inclusive, so neighbouring squares overlap by one pixel and the order counts."""
    grid = np.asarray(grid)
    H, W = grid.shape
    cs = cell_size
    ops = []
    kind = grid % 10
    for r, c in zip(*np.nonzero((grid == 1) | (grid == 2))):
        x, y = int(c) * cs, int(r) * cs
        if kind[r, c] == 1:
            ops.append(('rect', (x, y, x + cs, y + cs), COLOR_WALL))
        else:
            pad = cs * 0.2
            ops.append(('oval', (x + pad, y + pad, x + cs - pad, y + cs - pad), COLOR_FRUIT))
    for i in range(len(alive)):
        if not alive[i]:
            continue
        colour = SNAKE_COLORS[i % len(SNAKE_COLORS)]
        body = [int(v) for v in cells[i] if v >= 0]
        for cell in body:
            x, y = (cell % W) * cs, (cell // W) * cs
            ops.append(('rect', (x, y, x + cs, y + cs), colour))
        hx, hy = (body[0] % W) * cs, (body[0] // W) * cs
        ops.append(('oval', (hx, hy, hx + cs, hy + cs), colour))
        dy, dx = DELTA[int(dirs[i])]
        mx, my = hx + cs / 2, hy + cs / 2
        ahead, beside, eye = cs * 0.3, cs * 0.15, cs * 0.1
        centres = [(mx + dx * ahead - dy * beside, my + dy * ahead - dx * beside),
                   (mx + dx * ahead + dy * beside, my + dy * ahead + dx * beside)]
        for radius, colour2 in ((eye, (255, 255, 255)), (eye * 0.5, (0, 0, 0))):
            for ex, ey in centres:
                ops.append(('oval', (ex - radius, ey - radius, ex + radius, ey + radius), colour2))
    return ops


def render_fancy(grid, cells, alive, dirs, cell_size=40):
    """SnakeEnv.render_fancy: uint8 [H*cell_size, W*cell_size, 3].  cells: int [ns, L] head-first flat cell indices
    (-1 padded), alive / dirs: [ns]."""
    from PIL import Image, ImageDraw
    grid = np.asarray(grid)
    H, W = grid.shape
    image = Image.new('RGB', (W * cell_size, H * cell_size), COLOR_BG)
    draw = ImageDraw.Draw(image)
    for shape, box, colour in fancy_ops(grid, cells, alive, dirs, cell_size):
        (draw.rectangle if shape == 'rect' else draw.ellipse)(list(box), fill=colour)
    return np.array(image)
