"""TEST INFRASTRUCTURE ONLY -- a stand-in ``matplotlib`` so the reference imports unmodified.

The reference scripts do ``from matplotlib import animation, mlab`` at module top
(/root/reference/particle_filter.py:12, mylib/transform.py:10, mylib/error_ellipse.py:12-13,
motion_model.py:12).  matplotlib is not installed in this image, and ``mlab.bivariate_normal``
(the only arithmetic the particle filter takes from it, particle_filter.py:191) was removed in
matplotlib 3.1 anyway.  ``install()`` registers empty ``animation``/``patches``/``pyplot``
modules and an ``mlab`` module whose ``bivariate_normal`` restates the published
matplotlib <= 3.0 formula (mlab.py, ``bivariate_normal``):

    Xmu = X-mux; Ymu = Y-muy
    rho = sigmaxy/(sigmax*sigmay)
    z = Xmu**2/sigmax**2 + Ymu**2/sigmay**2 - 2*rho*Xmu*Ymu/(sigmax*sigmay)
    denom = 2*np.pi*sigmax*sigmay*np.sqrt(1-rho**2)
    return np.exp(-z/(2*(1-rho**2))) / denom

Only ``tests/golden/make_golden.py`` (run in the authoring container, where /root/reference
exists) uses this.  Nothing in the product path imports it.
"""
import sys
import types

import numpy as np


def bivariate_normal(X, Y, sigmax=1.0, sigmay=1.0, mux=0.0, muy=0.0, sigmaxy=0.0):
    Xmu = X - mux
    Ymu = Y - muy
    rho = sigmaxy / (sigmax * sigmay)
    z = Xmu ** 2 / sigmax ** 2 + Ymu ** 2 / sigmay ** 2 - 2 * rho * Xmu * Ymu / (sigmax * sigmay)
    denom = 2 * np.pi * sigmax * sigmay * np.sqrt(1 - rho ** 2)
    return np.exp(-z / (2 * (1 - rho ** 2))) / denom


def install():
    """Put the stub into sys.modules (no-op when a real matplotlib with mlab.bivariate_normal exists)."""
    if "matplotlib" in sys.modules and hasattr(
        getattr(sys.modules["matplotlib"], "mlab", None), "bivariate_normal"
    ):
        return
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []  # mark as package
    for name in ("animation", "patches", "pyplot", "mlab"):
        sub = types.ModuleType("matplotlib." + name)
        setattr(mpl, name, sub)
        sys.modules["matplotlib." + name] = sub
    mpl.mlab.bivariate_normal = bivariate_normal
    sys.modules["matplotlib"] = mpl
