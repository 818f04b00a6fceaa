"""Parity gates of BASELINE.json's north_star, shared by the CPU (emulator) and GPU tests.

* integer peak (round(dy), round(dx)) bit-exact, except DOCUMENTED TIES: the value of the
  oracle's correlation surface at the other peak location is within ``tie_rel`` (relative to the
  surface maximum) of the maximum -- i.e. the two candidates differ by float32 FFT rounding only;
* sub-pixel displacement within 0.02 px;
* confidence within 1e-4 relative, with a 1e-6 absolute floor for conf -> 0.
"""
import numpy as np

from oracle import xcorr_oracle as xo

SUBPIXEL_TOL = 0.02
CONF_RTOL = 1e-4
CONF_ATOL = 1e-6
TIE_REL = 1e-5


def _wrapdiff(a, b, period):
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    return d - np.round(d / period) * period


def check_against_oracle(got, img0, img1, conf_rtol=CONF_RTOL, conf_atol=CONF_ATOL, **kw):
    """``got`` = (dx, dy, conf) from the implementation under test.  Returns the number of
    documented ties encountered (peaks that differ but tie within float rounding)."""
    dx, dy, cf = (np.asarray(v) for v in got[:3])
    rdx, rdy, rcf, dbg = xo.xcorr_oracle(img0, img1, return_debug=True, **kw)
    return compare(dx, dy, cf, rdx, rdy, rcf, img0, img1, dbg['fftshp'], conf_rtol, conf_atol, **kw)


def compare(dx, dy, cf, rdx, rdy, rcf, img0=None, img1=None, fftshp=None, conf_rtol=CONF_RTOL, conf_atol=CONF_ATOL, **kw):
    assert dx.shape == rdx.shape and dy.shape == rdy.shape and cf.shape == rcf.shape
    assert dx.dtype == rdx.dtype and dy.dtype == rdy.dtype, (dx.dtype, rdx.dtype)
    assert cf.dtype == rcf.dtype, (cf.dtype, rcf.dtype)
    ties = 0
    if fftshp is None:
        a, b = np.asarray(img0), np.asarray(img1)
        fftshp = xo.fft_shape(a.shape[1:3], b.shape[1:3], kw.get('pad', True))
    ny, nx = fftshp
    ex = np.abs(_wrapdiff(dx, rdx, nx))
    ey = np.abs(_wrapdiff(dy, rdy, ny))
    bad = np.nonzero((ex > SUBPIXEL_TOL) | (ey > SUBPIXEL_TOL))[0]
    for i in bad:
        # a different integer peak is only acceptable as a documented tie
        assert img0 is not None, f'pair {i}: displacement differs ({dx[i]}, {dy[i]}) vs ({rdx[i]}, {rdy[i]})'
        surf, _, _ = xo.correlation_surfaces(np.asarray(img0)[i:i + 1], np.asarray(img1)[i:i + 1],
                                             pad=kw.get('pad', True), want_mirror=False)
        surf = surf[0]
        h0, w0 = np.asarray(img0).shape[1:3]
        h1, w1 = np.asarray(img1).shape[1:3]
        py = int(np.round(dy[i] - (h0 - h1) / 2)) % ny
        px = int(np.round(dx[i] - (w0 - w1) / 2)) % nx
        top = surf.max()
        assert top - surf[py, px] <= TIE_REL * max(abs(top), 1e-30), \
            f'pair {i}: peak differs and is not a tie: got ({dx[i]}, {dy[i]}), oracle ({rdx[i]}, {rdy[i]})'
        ties += 1
    ok = np.ones(dx.shape, bool)
    ok[bad] = False
    both_nan = np.isnan(cf) & np.isnan(rcf)
    sel = ok & ~both_nan
    err = np.abs(cf[sel].astype(np.float64) - rcf[sel].astype(np.float64))
    lim = conf_rtol * np.abs(rcf[sel].astype(np.float64)) + conf_atol
    assert np.all(err <= lim), f'conf mismatch: max err {err.max():.3e} (limit {lim[np.argmax(err - lim)]:.3e})'
    assert np.array_equal(np.isnan(cf), np.isnan(rcf))
    return ties
