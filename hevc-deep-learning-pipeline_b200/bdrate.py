"""Bjontegaard delta metrics (BD-rate, BD-PSNR) -- the evaluation step the reference leaves to the zipped
third-party tools under calc_BDBR/ (README.md:23, calc_BDBR/README.md:1-39).

Polynomial fit of PSNR over log10(rate) (BD-PSNR) or of log10(rate) over PSNR (BD-rate), integrated over the
common interval; order 3 least squares for four or more points (VCEG-M33, as Bjontegaard-python3.zip does), or
`order=len-1` (exact interpolation) as the five-point JCTVC-B055 tool does.  Pinned in
tests/test_bdrate.py against the worked example shipped in calc_BDBR/JCTVC-B055.zip
(reference.txt / proposal.txt -> 1.628122 dB, -35.976930 %).
"""
import numpy as np


def _fit_integrate(x, y, lo, hi, order):
    # centre x for conditioning; the integral is translation invariant
    c = 0.5 * (lo + hi)
    p = np.polyfit(np.asarray(x, np.float64) - c, np.asarray(y, np.float64), order)
    P = np.polyint(p)
    return np.polyval(P, hi - c) - np.polyval(P, lo - c)


def bd_psnr(rate_anchor, psnr_anchor, rate_test, psnr_test, order=3):
    """Average PSNR difference (dB) test - anchor over the common log-rate interval."""
    la, lt = np.log10(rate_anchor), np.log10(rate_test)
    order = min(order, len(la) - 1, len(lt) - 1)
    lo, hi = max(la.min(), lt.min()), min(la.max(), lt.max())
    return float((_fit_integrate(lt, psnr_test, lo, hi, order) - _fit_integrate(la, psnr_anchor, lo, hi, order)) / (hi - lo))


def bd_rate(rate_anchor, psnr_anchor, rate_test, psnr_test, order=3):
    """Average bitrate difference (%) test vs anchor at equal PSNR over the common PSNR interval."""
    la, lt = np.log10(rate_anchor), np.log10(rate_test)
    pa, pt = np.asarray(psnr_anchor, np.float64), np.asarray(psnr_test, np.float64)
    order = min(order, len(la) - 1, len(lt) - 1)
    lo, hi = max(pa.min(), pt.min()), min(pa.max(), pt.max())
    d = (_fit_integrate(pt, lt, lo, hi, order) - _fit_integrate(pa, la, lo, hi, order)) / (hi - lo)
    return float((10.0 ** d - 1.0) * 100.0)
