"""Synthetic inputs shared by the measurement tools (product classes and numpy only - the tools never touch oracle/)."""
import numpy as np


def mw3_product():
    """MW3 = Hernquist + Miyamoto-Nagai + NFW with the reference's GalaMilkyWayPotential values (potential.py:394-408)."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    return P.Potential_Combine([P.HernquistPotential(m=5e9, r_s=1.0, units=ssc.usys), P.MiyamotoNagaiDisk(m=6.8e10, a=3.0, b=0.28, units=ssc.usys),
                                P.NFWPotential(m=5.4e11, r_s=15.62, units=ssc.usys)], units=ssc.usys)


def halo_orbits(n, seed=0):
    """Stream-progenitor-like orbits: r in [12, 30] kpc, mostly tangential velocities."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    r = rng.uniform(12.0, 30.0, n)
    x = d * r[:, None]
    e = np.cross(d, rng.normal(size=(n, 3))); e /= np.linalg.norm(e, axis=1)[:, None]
    v = e * rng.uniform(0.14, 0.22, n)[:, None] + d * rng.normal(size=(n, 1)) * 0.03
    return np.hstack([x, v])


def subhalo_set(n, seed=1, t_lo=-3000.0, t_hi=0.0, tw=150.0):
    """Subhalo impacts in the style of generate_derivs.py:155: M ~ 10^U(5,9), r_s = 1.05 sqrt(M / 1e8), v ~ N(0, 0.184^2)."""
    rng = np.random.default_rng(seed)
    M = 10 ** rng.uniform(5, 9, n)
    rs = 1.05 * np.sqrt(M / 1e8)
    return dict(m=np.ones(n), M=M, rs=rs, x0=rng.normal(size=(n, 3)) * 10.0, v=rng.normal(size=(n, 3)) * 0.184, t0=rng.uniform(t_lo, t_hi, n),
                tw=np.full(n, tw))
