"""Match-list layouts against the reference's own reader / writer statements
(feabas/stitcher.py:144-151,194-208; feabas/aligner.py:26-44,134-141), restated inline."""
import numpy as np
import pytest

from feabas_b200.cuda import matchio


def test_stitch_payload_round_trip_and_reference_reader():
    rng = np.random.default_rng(0)
    xy0, xy1 = rng.uniform(0, 3000, (17, 2)), rng.uniform(0, 3000, (17, 2))
    w = rng.uniform(0, 1, 17).astype(np.float32)
    data = matchio.pack_stitch_match(xy0, xy1, w, 0.0123)
    assert data.dtype == np.float32 and data.shape == (17 * 5 + 1,)
    # the reference's reader, verbatim arithmetic (stitcher.py:201-206)
    npt = int((data.size - 1) / 5)
    np.testing.assert_array_equal(data[0:(2 * npt)].reshape(-1, 2), xy0.astype(np.float32))
    np.testing.assert_array_equal(data[(2 * npt):(4 * npt)].reshape(-1, 2), xy1.astype(np.float32))
    np.testing.assert_array_equal(data[(4 * npt):(5 * npt)], w)
    assert data[-1] == np.float32(0.0123)
    a, b, c, s = matchio.unpack_stitch_match(data)
    np.testing.assert_array_equal(a, xy0.astype(np.float32))
    np.testing.assert_array_equal(b, xy1.astype(np.float32))
    np.testing.assert_array_equal(c, w)
    assert s == np.float32(0.0123)
    with pytest.raises(ValueError):
        matchio.unpack_stitch_match(np.zeros(7, np.float32))
    with pytest.raises(ValueError):
        matchio.pack_stitch_match(xy0, xy1[:5], w, 0.0)
    e = matchio.pack_stitch_match(np.empty((0, 2)), np.empty((0, 2)), np.empty(0), 0.05)
    assert e.shape == (1,) and matchio.unpack_stitch_match(e)[0].shape == (0, 2)


def test_align_datasets_and_resolution_rescale():
    xy0 = np.array([[10.0, 20.0], [30.5, 40.25]])
    xy1 = xy0 + 1.5
    ds = matchio.align_match_datasets(xy0, xy1, np.array([[0.5], [0.75]]), 16.0, np.array(0.02), 'sec_0001', 'sec_0002')
    assert matchio.numpy_to_str_ascii(ds['name0']) == 'sec_0001' and ds['name1'].dtype == np.uint8
    m = matchio.match_from_datasets(ds)
    assert m.weight.shape == (2,) and m.strain == 0.02
    np.testing.assert_array_equal(m.xy0, xy0)
    m2 = matchio.match_from_datasets(ds, target_resolution=4.0)          # scale 4: (x + 0.5) * 4 - 0.5
    np.testing.assert_allclose(m2.xy0, 4.0 * (xy0 + 0.5) - 0.5)
    np.testing.assert_allclose(m2.xy1, 4.0 * (xy1 + 0.5) - 0.5)
    del ds['strain']
    assert matchio.match_from_datasets(ds).strain == 0.05                 # config.DEFAULT_AVG_DEFORM


def test_h5_helpers_need_h5py():
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            matchio.write_align_match('/tmp/x.h5', np.zeros((1, 2)), np.zeros((1, 2)), np.ones(1), 4.0, 0.05, 'a', 'b')
    else:                                                                 # pragma: no cover - h5py is absent here
        import os
        import tempfile
        path = os.path.join(tempfile.mkdtemp(), 'm.h5')
        matchio.write_align_match(path, np.zeros((3, 2)), np.ones((3, 2)), np.ones(3), 4.0, 0.05, 'a', 'b')
        m = matchio.read_align_match(path)
        assert m.xy1.shape == (3, 2) and m.strain == 0.05
