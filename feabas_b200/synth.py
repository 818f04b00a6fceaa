"""Seeded synthetic EM-like imagery for tests and benchmarks.

"EM-like" = band-limited texture: white Gaussian noise blurred with a small
Gaussian, plus sparse dark blobs, scaled to uint8 with mean ~128 / std ~40.
Pairs are cut from one canvas at known offsets with independent sensor noise,
so the true displacement of every pair is known.  Pure numpy (host side); the
benchmark uploads the result once.
"""
import numpy as np
from scipy.ndimage import gaussian_filter


def em_canvas(height, width, seed, blur=2.0, blobs_per_mpx=150):
    rng = np.random.default_rng(seed)
    tex = gaussian_filter(rng.standard_normal((height, width)).astype(np.float32), blur)
    tex /= tex.std() + 1e-12
    nblob = int(blobs_per_mpx * height * width / 1e6)
    if nblob:
        dots = np.zeros((height, width), dtype=np.float32)
        dots[rng.integers(0, height, nblob), rng.integers(0, width, nblob)] = 1.0
        dots = gaussian_filter(dots, 4.0)
        dots /= dots.max() + 1e-12
        tex -= 3.0 * dots
    img = 128.0 + 40.0 * tex
    return np.clip(img, 0, 255).astype(np.uint8)


def dog_f32(img_u8, sigma=2.5):
    """Cheap host band-pass used only to make float32 block stacks that look like
    the matcher's input; NOT the reference filter (see feabas_b200.common)."""
    f = img_u8.astype(np.float32)
    a = gaussian_filter(f, sigma, mode='nearest')
    return a - gaussian_filter(a, sigma, mode='nearest')


def block_pairs(n, size, seed, max_shift=32, noise=10.0, dtype=np.float32, band_pass=True):
    """``n`` pairs of ``size x size`` blocks with known integer displacement.

    Returns ``(stack0, stack1, shifts)`` where ``shifts[i] = (dx, dy)`` is the
    displacement ``xcorr_fft`` should report: content at ``p`` in ``stack0[i]``
    sits at ``p + (dx, dy)`` in ``stack1[i]``.
    """
    rng = np.random.default_rng(seed)
    size_h, size_w = (size, size) if np.isscalar(size) else size
    pad = max_shift + 8
    ch, cw = size_h + 2 * pad, size_w + 2 * pad
    per_row = max(1, int(np.ceil(np.sqrt(n))))
    canvas = em_canvas(ch * per_row, cw * per_row, seed)
    s0 = np.empty((n, size_h, size_w), dtype=dtype)
    s1 = np.empty((n, size_h, size_w), dtype=dtype)
    shifts = rng.integers(-max_shift, max_shift + 1, size=(n, 2))
    for i in range(n):
        oy, ox = (i // per_row) * ch + pad, (i % per_row) * cw + pad
        dx, dy = shifts[i]
        a = canvas[oy - pad:oy + size_h + pad, ox - pad:ox + size_w + pad].astype(np.float32)
        b = a + rng.normal(0, noise, a.shape).astype(np.float32)
        a = a + rng.normal(0, noise, a.shape).astype(np.float32)
        if band_pass:
            a, b = dog_f32(a), dog_f32(b)
        s0[i] = a[pad:pad + size_h, pad:pad + size_w].astype(dtype)
        s1[i] = b[pad - dy:pad - dy + size_h, pad - dx:pad - dx + size_w].astype(dtype)
    return s0, s1, shifts


def tile_grid(rows, cols, tile_hw=(3000, 4000), overlap=0.1, jitter=10, seed=1):
    """A ``rows x cols`` montage of overlapping uint8 tiles cut from one canvas.

    Returns ``(tiles, nominal_xy, true_xy)``: the tile images, the nominal
    stage positions (what a coordinate file would say) and the true positions
    (nominal + jitter) used to cut them.
    """
    rng = np.random.default_rng(seed)
    th, tw = tile_hw
    sy, sx = int(round(th * (1 - overlap))), int(round(tw * (1 - overlap)))
    margin = jitter + 2
    canvas = em_canvas(sy * (rows - 1) + th + 2 * margin, sx * (cols - 1) + tw + 2 * margin, seed)
    tiles, nominal, true = [], [], []
    for r in range(rows):
        for c in range(cols):
            jx, jy = rng.integers(-jitter, jitter + 1, size=2)
            x, y = c * sx, r * sy
            nominal.append((x, y))
            true.append((x + jx, y + jy))
            yy, xx = y + jy + margin, x + jx + margin
            tiles.append(np.ascontiguousarray(canvas[yy:yy + th, xx:xx + tw]))
    return tiles, np.array(nominal), np.array(true)
