"""Run the UNMODIFIED reference matcher layers in the build container and record what they do.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Only usable where ``/root/reference`` exists; used by
``oracle/make_golden.py`` to write ``tests/golden/loop_*.npz``.

What executes unmodified (imported from ``/root/reference`` by ``oracle/ref_loader.py``):

* ``feabas.matcher.stitching_matcher``               matcher.py:224-367
* ``feabas.matcher.section_matcher``                 matcher.py:370-427
* ``feabas.matcher.iterative_xcorr_matcher_w_mesh``  matcher.py:430-778
* ``feabas.matcher.bboxes_mesh_renderer_matcher``    matcher.py:781-861
* ``feabas.matcher.global_translation_matcher`` / ``xcorr_fft`` / ``distributor_cartesian_bbox``
* ``feabas.renderer.MeshRenderer.crop_multiple`` / ``crop_field`` / ``crop_field_affine``  renderer.py:419-648
* ``feabas.common.render_by_subregions`` / ``remap`` / ``masked_dog_filter`` / ``divide_bbox`` ...
* ``feabas.dal.StreamLoader``                        dal.py:1008-1050
* ``feabas.spatial.scale_coordinates``, ``feabas.config.data_resolution``

What is replaced -- the geometry layer only (its dependencies ``triangle``,
``shapely``, ``matplotlib``, ``pyamg`` are not installed here):

* ``matcher.Mesh``            -> ``AffineMesh`` (one affine map per section; ``from_bbox`` puts the corner vertices
                                 at ``bbox - 0.5`` exactly as ``Mesh.from_bbox(cartesian=True)``, mesh.py:426-427)
* ``matcher.optimizer.SLM``   -> ``AffineSLM`` (weighted least-squares affine relaxation)
* ``matcher.MeshRenderer``    -> a SUBCLASS of the reference's ``MeshRenderer`` whose only override is the
                                 constructor ``from_mesh``: it fills in the ``global_affine`` approximator
                                 (residue 0) and the covered region, so that the reference's own ``crop_field``
                                 takes its ``crop_field_affine`` branch (renderer.py:501-507)
* the five ``shapely`` calls of ``crop_field_affine`` (``box``, ``affine_transform``, ``intersection``, ``area``,
  ``contains_xy``) -> a convex-polygon implementation below (the region of an ``AffineMesh`` is a rectangle)

``AffineMesh`` / ``AffineSLM`` are the stand-ins the product ships for FEABAS-less use
(``feabas_b200/cuda/surrogate.py``): both sides of the comparison relax with the same model, everything
pixel-related on the reference side is the reference's own code.
"""
import types

import numpy as np

from . import ref_loader
from .convex import box as _box, affine_transform as _affine_transform, contains_xy as _contains_xy


# --------------------------------------------------------------------------------------------
# the harness
# --------------------------------------------------------------------------------------------
class Harness:
    """``h = Harness()``; ``h.matcher`` is the reference module with the geometry layer swapped; every call of
    ``xcorr_fft`` and ``bboxes_mesh_renderer_matcher`` made through it is appended to ``h.trace``."""

    def __init__(self):
        from feabas_b200.cuda.surrogate import AffineMesh, AffineSLM
        matcher, common, const = ref_loader.load()
        import importlib
        renderer = importlib.import_module('feabas.renderer')
        dal = importlib.import_module('feabas.dal')
        self.matcher, self.common, self.const, self.renderer, self.dal = matcher, common, const, renderer, dal
        self.AffineMesh, self.AffineSLM = AffineMesh, AffineSLM
        self.trace = []
        self._saved = {}

        ref_renderer_cls = renderer.MeshRenderer

        class AffineMeshRenderer(ref_renderer_cls):
            """The reference renderer over an ``AffineMesh``: only the constructor differs."""

            @classmethod
            def from_mesh(cls, srcmesh, gear=(const.MESH_GEAR_MOVING, const.MESH_GEAR_INITIAL), **kwargs):
                tol = kwargs.pop('affine_approx_tol', 0)
                kwargs.pop('weight_params', None)
                ainv, tinv = srcmesh.render_map(gear[0])              # p_initial = p_moving @ ainv + tinv
                full = np.concatenate((ainv, tinv.reshape(1, 2)), axis=0)
                approx = {'global_affine': full, 'global_residue': 0.0}
                x0, y0, x1, y1 = srcmesh.bounds                       # vertex extents, INITIAL gear
                covered = _box(x0 + 0.5, y0 + 0.5, x1 - 0.5, y1 - 0.5)   # shapely_regions(...).buffer(-0.5), renderer.py:100
                kwargs.pop('cache', None)
                # crop_field only looks at the approximator when the tolerance is positive (renderer.py:499); an
                # affine mesh is reproduced exactly by its global affine map, so any positive value selects it
                return cls([None], offset=np.zeros((1, 2)), resolution=srcmesh.resolution, affine_approximator=approx,
                           affine_approx_tol=max(float(tol), 1e-9), covered_region=covered, **kwargs)

        self.renderer_cls = AffineMeshRenderer

    # -- patching ----------------------------------------------------------------------------
    def __enter__(self):
        m, r = self.matcher, self.renderer
        self._saved = dict(Mesh=m.Mesh, optimizer=m.optimizer, MeshRenderer=m.MeshRenderer, xcorr_fft=m.xcorr_fft,
                           bboxes=m.bboxes_mesh_renderer_matcher, shpgeo=r.shpgeo, shapely=r.shapely)
        m.Mesh = self.AffineMesh
        m.optimizer = types.SimpleNamespace(SLM=self.AffineSLM)
        m.MeshRenderer = self.renderer_cls
        r.shpgeo = types.SimpleNamespace(box=_box)
        r.shapely = types.SimpleNamespace(affinity=types.SimpleNamespace(affine_transform=_affine_transform),
                                          prepare=lambda g: None, contains_xy=_contains_xy)
        ref_xcorr, ref_bboxes, trace = self._saved['xcorr_fft'], self._saved['bboxes'], self.trace

        def xcorr_logged(img0, img1, conf_mode=self.const.FFT_CONF_MIRROR, **kwargs):
            out = ref_xcorr(img0, img1, conf_mode=conf_mode, **kwargs)
            trace.append(dict(kind='xcorr', shape0=np.array(np.shape(img0)), shape1=np.array(np.shape(img1)),
                              pad=bool(kwargs.get('pad', True)), subpixel=bool(kwargs.get('subpixel', False)),
                              dx=out[0].copy(), dy=out[1].copy(), conf=out[2].copy()))
            return out

        def bboxes_logged(mesh0, mesh1, loader0, loader1, bboxes0, bboxes1, **kwargs):
            out = ref_bboxes(mesh0, mesh1, loader0, loader1, bboxes0, bboxes1, **kwargs)
            trace.append(dict(kind='level', bboxes0=np.array(bboxes0, dtype=np.float64), bboxes1=np.array(bboxes1, dtype=np.float64),
                              pad=bool(kwargs.get('pad', True)), subpixel=bool(kwargs.get('subpixel', False)),
                              batch_size=-1 if kwargs.get('batch_size', None) is None else int(kwargs['batch_size']),
                              sigma=float(kwargs.get('sigma', 0.0)),
                              map0=np.concatenate([np.ravel(v) for v in mesh0.get_map()]),
                              map1=np.concatenate([np.ravel(v) for v in mesh1.get_map()]),
                              xy0=np.array(out[0]), xy1=np.array(out[1]), conf=np.array(out[2])))
            return out

        m.xcorr_fft = xcorr_logged
        m.bboxes_mesh_renderer_matcher = bboxes_logged
        return self

    def __exit__(self, *exc):
        m, r, s = self.matcher, self.renderer, self._saved
        m.Mesh, m.optimizer, m.MeshRenderer = s['Mesh'], s['optimizer'], s['MeshRenderer']
        m.xcorr_fft, m.bboxes_mesh_renderer_matcher = s['xcorr_fft'], s['bboxes']
        r.shpgeo, r.shapely = s['shpgeo'], s['shapely']
        return False

    def take_trace(self):
        t, self.trace[:] = list(self.trace), []
        return t

    def stream_loader(self, img, **kwargs):
        return self.dal.StreamLoader(img, **kwargs)


def flatten_trace(prefix, trace, blob):
    """Store a trace (list of dicts of arrays / scalars) under ``prefix/<i>/<key>`` of a flat npz dict."""
    blob[f'{prefix}/n'] = np.asarray(len(trace))
    for i, rec in enumerate(trace):
        for k, v in rec.items():
            blob[f'{prefix}/{i}/{k}'] = np.asarray(v)
