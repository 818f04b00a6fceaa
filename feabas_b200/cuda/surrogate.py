"""Affine stand-ins for the reference's geometry layer, so that the matcher control loop can run
without FEABAS' ``Mesh`` / ``SLM`` / ``MeshRenderer`` (triangle, shapely, pyamg ... are not part
of this package and not available in the build container).

``AffineMesh`` carries ONE affine map per section instead of a triangulated elastic mesh, and
``AffineSLM`` "relaxes" it by a weighted least-squares affine fit of the current links (plus the
reference's Huber / threshold re-weighting of link residues).  They expose the same methods the
loop in ``matcher.py`` calls on the reference objects (feabas/matcher.py:541-751), so the loop
itself is backend agnostic.  Results obtained with them are a *surrogate loop*: block geometry,
level selection, confidence filtering and every ``xcorr_fft`` call are the reference's, the
relaxation between levels is not.

``ArrayLoader`` mimics ``dal.StreamLoader`` (feabas/dal.py:1008-1050) for an image that lives in
GPU memory.
"""
import numpy as np

from .constant import MESH_GEAR_FIXED, MESH_GEAR_INITIAL, MESH_GEAR_MOVING, MESH_GEAR_STAGING, DEFAULT_RESOLUTION

try:
    import torch
except Exception:                       # pragma: no cover
    torch = None


class ArrayLoader:
    """In-memory image on the GPU with the loader attributes the matcher reads."""

    def __init__(self, img, fillval=0, resolution=DEFAULT_RESOLUTION, x0=0, y0=0, device=None):
        from . import image as _img
        self._ready = None
        src = img if (torch is not None and isinstance(img, torch.Tensor)) else None
        if src is None and torch is not None and isinstance(img, np.ndarray) and img.flags.c_contiguous and torch.cuda.is_available():
            src = torch.from_numpy(img)
        if src is not None and not src.is_cuda and src.dim() == 2 and torch.cuda.is_available() and src.is_pinned():
            # page-locked host image: upload on a side stream, so that the copy of the NEXT section overlaps the
            # block extraction / band-pass of this one; consumers wait on the event at first use of .tensor
            dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
            up = _img.upload_stream(dev)
            with torch.cuda.stream(up):
                self._tensor = src.to(dev, non_blocking=True)
            self._ready = torch.cuda.Event()
            self._ready.record(up)
        else:
            self._tensor = _img.to_device(img, device)
        if self._tensor.dim() != 2:
            raise ValueError('ArrayLoader holds a single-channel 2-D image')
        self.default_fillval = fillval
        self.resolution = resolution
        self.x0, self.y0 = x0, y0

    @property
    def tensor(self):
        """The image on the GPU (the current stream is made to wait for a pending upload)."""
        if self._ready is not None:
            cur = torch.cuda.current_stream(self._tensor.device)
            cur.wait_event(self._ready)
            self._tensor.record_stream(cur)
            self._ready = None
        return self._tensor

    @property
    def dtype(self):
        return np.dtype(str(self._tensor.dtype).replace('torch.', ''))

    @property
    def bounds(self):
        h, w = self._tensor.shape
        return (self.x0, self.y0, self.x0 + w, self.y0 + h)

    def crop(self, bbox, return_empty=False, **kwargs):
        """Host copy of ``bbox`` = (xmin, ymin, xmax, ymax), ``fillval`` outside the image."""
        fill = kwargs.get('fillval', self.default_fillval)
        x_lo, y_lo, x_hi, y_hi = (int(v) for v in bbox)
        out = np.full((y_hi - y_lo, x_hi - x_lo), fill, dtype=self.dtype)
        h, w = self.tensor.shape
        sy0, sy1 = max(y_lo - self.y0, 0), min(y_hi - self.y0, h)
        sx0, sx1 = max(x_lo - self.x0, 0), min(x_hi - self.x0, w)
        if sy1 <= sy0 or sx1 <= sx0:
            return out if return_empty else None
        out[sy0 + self.y0 - y_lo:sy1 + self.y0 - y_lo, sx0 + self.x0 - x_lo:sx1 + self.x0 - x_lo] = \
            self.tensor[sy0:sy1, sx0:sx1].cpu().numpy()
        return out


_IDENTITY8 = None


def _rank_deficient(centered):
    """``np.linalg.matrix_rank(centered) < 2`` for an N x 2 array (same singular values, same tolerance rule)."""
    sv = np.linalg.svd(centered, compute_uv=False)
    return int(np.count_nonzero(sv > sv.max() * max(centered.shape) * np.finfo(np.float64).eps)) < 2


class AffineMesh:
    """A section whose deformation is one affine map per gear: ``p_gear = p_initial @ A + t``."""

    is_linear = True
    soft_factor = 1.0

    def __init__(self, bounds, uid=0, resolution=DEFAULT_RESOLUTION):
        self.bounds = tuple(float(v) for v in bounds)
        self.uid = uid
        self.resolution = resolution
        self.locked = False
        eye = (np.eye(2), np.zeros(2))
        self._maps = {MESH_GEAR_INITIAL: eye, MESH_GEAR_FIXED: eye, MESH_GEAR_MOVING: eye, MESH_GEAR_STAGING: eye}

    @classmethod
    def from_bbox(cls, bbox, cartesian=False, **kwargs):
        """Same call as ``Mesh.from_bbox`` (feabas/mesh.py:403-437).  The reference puts the outer vertices of the
        grid at ``bbox - 0.5`` (mesh.py:426-427): the mesh covers the AREA of the pixels ``xmin .. xmax - 1`` whose
        centres sit at integer coordinates.  ``mesh_size`` / ``min_num_blocks`` (the grid's density) have no
        meaning for a single affine map and are ignored."""
        x0, y0, x1, y1 = (float(v) for v in bbox)
        return cls((x0 - 0.5, y0 - 0.5, x1 - 0.5, y1 - 0.5), uid=kwargs.get('uid', 0),
                   resolution=kwargs.get('resolution', DEFAULT_RESOLUTION))

    def covered_rect(self):
        """(xmin, ymin, xmax, ymax), INITIAL gear: the mesh region shrunk by half a pixel, i.e. what
        ``MeshRenderer.from_mesh`` keeps as ``covered_region`` (feabas/renderer.py:98-101); only source positions
        STRICTLY inside it count as covered (``shapely.contains_xy``, renderer.py:447)."""
        x0, y0, x1, y1 = self.bounds
        return (x0 + 0.5, y0 + 0.5, x1 - 0.5, y1 - 0.5)

    # -- state -------------------------------------------------------------------------------
    def copy(self):
        other = AffineMesh(self.bounds, uid=self.uid, resolution=self.resolution)
        other.locked = self.locked
        other._maps = {g: (a.copy(), t.copy()) for g, (a, t) in self._maps.items()}
        return other

    def lock(self):
        self.locked = True

    def unlock(self):
        self.locked = False

    def get_map(self, gear=MESH_GEAR_MOVING):
        return self._maps[gear]

    def set_map(self, a, t, gear=MESH_GEAR_MOVING):
        if gear == MESH_GEAR_INITIAL:
            raise ValueError('the initial gear is read-only')
        self._maps[gear] = (np.array(a, dtype=np.float64).reshape(2, 2), np.array(t, dtype=np.float64).reshape(2))

    def apply_translation(self, dxy, gear=MESH_GEAR_FIXED, **kwargs):
        """Shift the vertices of ``gear`` (and of the gears derived from it, as ``Mesh.apply_translation``
        followed by the anneal at the top of the matching loop would)."""
        if self.locked:
            return
        dxy = np.asarray(dxy, dtype=np.float64).reshape(2)
        gears = (MESH_GEAR_FIXED, MESH_GEAR_MOVING, MESH_GEAR_STAGING) if gear == MESH_GEAR_FIXED else (gear,)
        for g in gears:
            a, t = self._maps[g]
            self._maps[g] = (a, t + dxy)

    def anneal(self, gear=(MESH_GEAR_MOVING, MESH_GEAR_FIXED), mode=None):
        """Copy the state of ``gear[0]`` into ``gear[1]`` (a locked mesh does not move, feabas/mesh.py:2426)."""
        if self.locked:
            return
        a, t = self._maps[gear[0]]
        if gear[1] != MESH_GEAR_INITIAL:
            self._maps[gear[1]] = (a.copy(), t.copy())

    # -- geometry ----------------------------------------------------------------------------
    def corners(self, gear=MESH_GEAR_MOVING):
        x0, y0, x1, y1 = self.bounds
        return self.transform(np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]]), gear)

    def vertices(self, gear=MESH_GEAR_MOVING):
        return self.corners(gear)

    def bbox(self, gear=MESH_GEAR_MOVING, **kwargs):
        c = self.corners(gear)
        return np.array([c[:, 0].min(), c[:, 1].min(), c[:, 0].max(), c[:, 1].max()])

    def transform(self, xy, gear=MESH_GEAR_MOVING):
        a, t = self._maps[gear]
        return np.asarray(xy, dtype=np.float64) @ a + t

    def inverse(self, xy, gear=MESH_GEAR_MOVING):
        a, t = self._maps[gear]
        return (np.asarray(xy, dtype=np.float64) - t) @ np.linalg.inv(a)

    def render_map(self, gear=MESH_GEAR_MOVING):
        """(Ainv, tinv): ``p_initial = p_gear @ Ainv + tinv`` -- what a renderer samples the image with."""
        a, t = self._maps[gear]
        ainv = np.linalg.inv(a)
        return ainv, -t @ ainv

    def connected_triangles(self):
        return 1, None

    def triangle_mask_for_stiffness(self, **kwargs):
        """One homogeneous piece of default material: nothing is soft (feabas/matcher.py:386-390)."""
        return np.ones(1, dtype=bool)

    def triangle_mask_for_render(self, **kwargs):
        return np.ones(1, dtype=bool)

    def submesh(self, tri_mask, **kwargs):
        return self if np.all(tri_mask) else None

    def stiffness_matrix(self, **kwargs):
        """Identity "stiffness" over the four corner vertices: with it the strain measure of the matcher is the
        RMS corner displacement relative to the RMS corner radius."""
        global _IDENTITY8
        if _IDENTITY8 is None:
            from scipy import sparse
            _IDENTITY8 = sparse.identity(8, format='csr')
        return _IDENTITY8, None


class _AffineLink:
    def __init__(self, mesh0, mesh1, xy0_initial, xy1_initial, weight):
        self.meshes = (mesh0, mesh1)
        self._xy0 = np.asarray(xy0_initial, dtype=np.float64).reshape(-1, 2)
        self._xy1 = np.asarray(xy1_initial, dtype=np.float64).reshape(-1, 2)
        self._weight = np.asarray(weight, dtype=np.float64).reshape(-1) if np.ndim(weight) else np.full(self._xy0.shape[0], float(weight))
        self._residue_weight = np.ones_like(self._weight)
        self._weight_func = None

    @property
    def mask(self):
        return (self._weight * self._residue_weight) > 0

    def _pick(self, arr, use_mask):
        return arr[self.mask] if use_mask else arr

    def xy0(self, gear=MESH_GEAR_MOVING, use_mask=False, combine=True):
        return self._pick(self.meshes[0].transform(self._xy0, gear), use_mask)

    def xy1(self, gear=MESH_GEAR_MOVING, use_mask=False, combine=True):
        return self._pick(self.meshes[1].transform(self._xy1, gear), use_mask)

    def weight(self, use_mask=False):
        return self._pick(self._weight * self._residue_weight, use_mask)

    def residues(self, gear=MESH_GEAR_MOVING):
        d = self.xy1(gear) - self.xy0(gear)
        return np.sum(d * d, axis=-1) ** 0.5


class AffineSLM:
    """Two ``AffineMesh`` sections tied by point links; relaxation = weighted affine least squares."""

    def __init__(self, meshes, stiffness_lambda=1.0, **kwargs):
        self.meshes = list(meshes)
        self.links = []
        self._stiffness_lambda = stiffness_lambda
        self._residue_rule = None

    def _mesh(self, uid):
        for m in self.meshes:
            if m.uid == uid:
                return m
        raise KeyError(uid)

    def clear_links(self):
        self.links = []

    def add_link_from_coordinates(self, uid0, uid1, xy0, xy1, gear=(MESH_GEAR_INITIAL, MESH_GEAR_INITIAL), weight=None, **kwargs):
        m0, m1 = self._mesh(uid0), self._mesh(uid1)
        xy0 = np.asarray(xy0, dtype=np.float64).reshape(-1, 2)
        xy1 = np.asarray(xy1, dtype=np.float64).reshape(-1, 2)
        if xy0.shape[0] == 0:
            return False
        if weight is None:
            weight = np.ones(xy0.shape[0])
        self.links.append(_AffineLink(m0, m1, m0.inverse(xy0, gear[0]), m1.inverse(xy1, gear[1]), weight))
        return True

    # -- relaxation --------------------------------------------------------------------------
    @staticmethod
    def _fit(src, dst, w, rigid=False):
        """Weighted least squares affine ``dst ~ src @ A + t`` (projected onto a rotation when ``rigid``); falls
        back to a translation for < 3 points."""
        a, t = AffineSLM._fit_affine(src, dst, w)
        if rigid:
            u, _, vh = np.linalg.svd(a)
            a = u @ vh
            w = np.asarray(w, dtype=np.float64)
            t = np.average(dst - src @ a, axis=0, weights=w) if w.sum() > 0 else np.zeros(2)
        return a, t

    @staticmethod
    def _fit_affine(src, dst, w):
        w = np.asarray(w, dtype=np.float64)
        if src.shape[0] < 3 or _rank_deficient(src - src.mean(0)):
            t = np.average(dst - src, axis=0, weights=w) if w.sum() > 0 else np.zeros(2)
            return np.eye(2), t
        sw = np.sqrt(w)[:, None]
        design = np.concatenate((src, np.ones((src.shape[0], 1))), axis=1)
        sol, *_ = np.linalg.lstsq(design * sw, dst * sw, rcond=None)
        return sol[:2], sol[2]

    def _gather(self, gear):
        p0 = np.concatenate([l.xy0(gear) for l in self.links], axis=0)
        p1 = np.concatenate([l.xy1(gear) for l in self.links], axis=0)
        w = np.concatenate([l.weight() for l in self.links], axis=0)
        return p0, p1, w

    def optimize_linear(self, **kwargs):
        """Move the unlocked section(s) so that the links close, in the MOVING gear."""
        if not self.links:
            return 0.0
        m0, m1 = self.links[-1].meshes
        p0, p1, w = self._gather(MESH_GEAR_MOVING)
        if not np.any(w > 0) or (m0.locked and m1.locked):
            return 0.0
        if m0.locked or not m1.locked:
            # bring mesh1 onto mesh0 (fully when mesh0 is locked, half way when both are free)
            a, t = self._fit(p1, p0, w)
            if not m0.locked:
                a, t = (a + np.eye(2)) / 2, t / 2
            a1, t1 = m1.get_map(MESH_GEAR_MOVING)
            m1.set_map(a1 @ a, t1 @ a + t, MESH_GEAR_MOVING)
        if m1.locked or not m0.locked:
            p0, p1, w = self._gather(MESH_GEAR_MOVING)
            a, t = self._fit(p0, p1, w)
            a0, t0 = m0.get_map(MESH_GEAR_MOVING)
            m0.set_map(a0 @ a, t0 @ a + t, MESH_GEAR_MOVING)
        return float(np.average(self.links[-1].residues(), weights=np.maximum(self.links[-1].weight(), 1e-12)))

    def optimize_Newton_Raphson(self, **kwargs):
        return self.optimize_linear(**kwargs)

    def optimize_affine_cascade(self, start_gear=MESH_GEAR_FIXED, target_gear=MESH_GEAR_FIXED, svd_clip=None, **kwargs):
        """Align the sections by one affine fit (rigid when ``svd_clip == (1, 1)``) of the links evaluated in
        ``start_gear``; result in ``target_gear``."""
        if not self.links:
            return
        rigid = svd_clip is not None and tuple(np.atleast_1d(svd_clip)) == (1, 1)
        m0, m1 = self.links[-1].meshes
        p0, p1, w = self._gather(start_gear)
        if m0.locked and m1.locked:
            return
        if m1.locked:
            a, t = self._fit(p0, p1, w, rigid)
            a0, t0 = m0.get_map(start_gear)
            m0.set_map(a0 @ a, t0 @ a + t, target_gear)
        else:
            a, t = self._fit(p1, p0, w, rigid)
            a1, t1 = m1.get_map(start_gear)
            m1.set_map(a1 @ a, t1 @ a + t, target_gear)
            if target_gear != start_gear:
                m0.anneal(gear=(start_gear, target_gear))

    def anneal(self, gear=(MESH_GEAR_FIXED, MESH_GEAR_MOVING), mode=None):
        for m in self.meshes:
            m.anneal(gear=gear, mode=mode)

    # -- residue driven re-weighting (feabas/optimizer.py: set_link_residue_huber / _threshold) ----
    def set_link_residue_huber(self, residue_len):
        self._residue_rule = ('huber', float(residue_len))

    def set_link_residue_threshold(self, residue_len):
        self._residue_rule = ('threshold', float(residue_len))

    def adjust_link_weight_by_residue(self, gear=MESH_GEAR_MOVING, relax_first=False):
        if self._residue_rule is None:
            return False, 0
        kind, length = self._residue_rule
        changed = False
        for link in self.links:
            r = link.residues(gear)
            if kind == 'huber':
                new = np.where(r > length, length / np.maximum(r, 1e-30), 1.0)
            else:
                new = (r <= length).astype(np.float64)
            if relax_first:
                new = np.maximum(new, 0.0)
            if np.any(np.abs(new - link._residue_weight) > 1e-3):
                changed = True
            link._residue_weight = new
        return changed, 0

    def divide_disconnected_submeshes(self, **kwargs):
        return 1
