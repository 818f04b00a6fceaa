"""Match-list containers either side of the matcher: the array layouts FEABAS stores
matches in, so that results of this package can be handed to ``stitch_main.py`` / ``align_main.py`` unchanged.

* stitching: one 1-D float32 dataset per overlap, ``concat(xy0, xy1, weight, strain)`` flattened
  (feabas/stitcher.py:144-151; reader :194-208: ``Npt = (size - 1) / 5``);
* alignment: datasets ``xy0``, ``xy1``, ``weight`` (gzip), ``resolution``, ``strain``, ``name0``, ``name1`` per section
  pair (feabas/aligner.py:134-141; reader :26-44).

The flat / dict forms are pure numpy.  The HDF5 files themselves need ``h5py`` (a dependency of FEABAS, not of this
package): the two file helpers raise ``ImportError`` when it is missing.
"""
import numpy as np


def pack_stitch_match(xy0, xy1, weight, strain):
    """``matches/<i>_<j>`` payload of a stitch H5 file (stitcher.py:144-151)."""
    xy0, xy1 = np.asarray(xy0), np.asarray(xy1)
    weight = np.asarray(weight)
    if xy0.shape != xy1.shape or xy0.ndim != 2 or xy0.shape[1] != 2 or weight.size != xy0.shape[0]:
        raise ValueError('xy0 / xy1 must be N x 2 and weight of length N')
    data = np.concatenate((xy0, xy1, weight, strain), axis=None)
    return data.astype(np.float32, copy=False)


def unpack_stitch_match(data):
    """Inverse of :func:`pack_stitch_match` (stitcher.py:194-208) -> ``(xy0, xy1, weight, strain)``."""
    data = np.asarray(data)
    if data.ndim != 1 or (data.size - 1) % 5:
        raise ValueError('not a stitch match payload: size %d is not 5 N + 1' % data.size)
    npt = int((data.size - 1) / 5)
    xy0 = data[0:(2 * npt)].reshape(-1, 2)
    xy1 = data[(2 * npt):(4 * npt)].reshape(-1, 2)
    weight = data[(4 * npt):(5 * npt)]
    return xy0, xy1, weight, data[-1]


def str_to_numpy_ascii(s):
    """feabas/common.py:438-440."""
    return np.frombuffer(s.encode('ascii'), dtype=np.uint8)


def numpy_to_str_ascii(ar):
    """feabas/common.py:433-435."""
    return np.asarray(ar).clip(0, 255).astype(np.uint8).ravel().tobytes().decode('ascii')


def align_match_datasets(xy0, xy1, weight, resolution, strain, name0, name1):
    """Dataset name -> array of one alignment match file (aligner.py:134-141)."""
    return {'xy0': np.asarray(xy0), 'xy1': np.asarray(xy1), 'weight': np.asarray(weight), 'resolution': resolution,
            'strain': strain, 'name0': str_to_numpy_ascii(name0), 'name1': str_to_numpy_ascii(name1)}


def match_from_datasets(ds, target_resolution=None, default_strain=0.05):
    """``read_matches_from_h5`` (aligner.py:26-44) on an already loaded mapping -> ``Match``."""
    from .constant import Match
    xy0, xy1 = np.asarray(ds['xy0']), np.asarray(ds['xy1'])
    weight = np.asarray(ds['weight']).ravel()
    resolution = ds['resolution']
    if isinstance(resolution, np.ndarray):
        resolution = resolution.item()
    strain = ds['strain'] if 'strain' in ds else default_strain
    if isinstance(strain, np.ndarray):
        strain = strain.item()
    if target_resolution is not None:
        scale = resolution / target_resolution
        xy0 = scale_coordinates(xy0, scale)
        xy1 = scale_coordinates(xy1, scale)
    return Match(xy0, xy1, weight, strain)


def scale_coordinates(coordinates, scale):
    """Pixel-centre preserving rescale (feabas/spatial.py:77-89)."""
    coordinates = np.array(coordinates, copy=False) if isinstance(coordinates, np.ndarray) else np.array(coordinates)
    if np.all(np.asarray(scale) == 1):
        return coordinates
    return scale * (coordinates + 0.5) - 0.5


def _h5py():
    try:
        import h5py
    except ImportError as exc:
        raise ImportError('writing / reading FEABAS match files needs h5py (a FEABAS dependency)') from exc
    return h5py


def write_align_match(path, xy0, xy1, weight, resolution, strain, name0, name1):
    h5py = _h5py()
    ds = align_match_datasets(xy0, xy1, weight, resolution, strain, name0, name1)
    with h5py.File(path, 'w') as f:
        for key in ('xy0', 'xy1', 'weight'):
            f.create_dataset(key, data=ds[key], compression='gzip')
        for key in ('resolution', 'strain', 'name0', 'name1'):
            f.create_dataset(key, data=ds[key])
    return len(ds['xy0'])


def read_align_match(path, target_resolution=None):
    h5py = _h5py()
    with h5py.File(path, 'r') as f:
        ds = {k: f[k][()] for k in f.keys()}
    return match_from_datasets(ds, target_resolution=target_resolution)
