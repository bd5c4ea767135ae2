"""Potential classes of the hot path - constructor signatures as in /root/reference/streamsculptor/potential.py.

Each class lowers itself to components of the flat potential program the CUDA kernels interpret
(include/ssb200.h).  Classes of the reference that wrap special functions, external libraries or arbitrary Python
callables (Zhao, PowerLawCutoff, AGAMA, SIDM, CustomPotential, ...) are out of scope for the B200 hot path
 and raise NotImplementedError instead of falling back to a CPU path.
"""
import os

import numpy as np

from . import _lib
from . import _runtime as rt
from .main import Potential
from .units import usys  # noqa: F401


# ---- tabulated tracks (what "center_spl" / "velocity_func" are on the device) ------------------------------------
def LinearTrack(t, y):
    """Piecewise-linear 3-vector track with linear extrapolation = jax RegularGridInterpolator(method='linear',
    bounds_error=False, fill_value=None) as used at potential.py:581-600."""
    return rt.Track(_lib.TRACK_LINEAR, t, y)


def CubicTrack(t, y):
    """C1 cubic Hermite track = interpax.Interpolator1D(method='cubic') as used at streamhelpers.py:520,
    perturbative.py:642 (NaN outside the knots)."""
    return rt.Track(_lib.TRACK_CUBIC, t, y)


def NotAKnotTrack(t, y):
    """C2 cubic spline through the knots with not-a-knot end conditions - what jax_cosmo's InterpolatedUnivariateSpline(k=3) builds, the spline the
    reference uses for the LMC orbit (potential.py:47-49) and the progenitor track of the restricted N-body driver (RestrictedNbody.py:34).  The
    spline is solved once on the host (scipy, O(n)); the kernels evaluate it exactly, as the cubic Hermite segments it consists of (knot values
    + knot slopes).  NaN outside the knots, like CubicTrack (the reference's spline extrapolates there)."""
    from scipy.interpolate import CubicSpline
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(len(t), 3)
    return rt.Track(_lib.TRACK_CUBIC, t, y, slopes=CubicSpline(t, y, axis=0, bc_type='not-a-knot')(t, 1))


def _no_nested(track):
    if track >= 0:
        raise NotImplementedError("nested time-dependent translations are not implemented")


class LMCPotential(Potential):                        # potential.py:40-63
    """NFW sphere (LMC_internal['m_NFW'], ['r_s_NFW']) translating along a cubic spline through LMC_orbit = {'x', 'y', 'z', 't'}."""

    def __init__(self, LMC_internal, LMC_orbit, units=None):
        super().__init__(units, {'LMC_internal': LMC_internal, 'LMC_orbit': LMC_orbit})
        self._track = NotAKnotTrack(LMC_orbit['t'], np.stack([np.asarray(LMC_orbit[k], dtype=np.float64) for k in ('x', 'y', 'z')], axis=1))

    def _lower(self, prog, track):
        _no_nested(track)
        from .units import usys, resolve_G
        G = resolve_G(usys)                            # potential.py:57: the inner NFW is built with units=usys, whatever `units` is
        prog.add(_lib.NFW, [G * self.LMC_internal['m_NFW'], self.LMC_internal['r_s_NFW']], prog.add_track(self._track))


class MiyamotoNagaiDisk(Potential):                   # potential.py:66-72
    def __init__(self, m, a, b, units=None):
        super().__init__(units, {'m': m, 'a': a, 'b': b})

    def _lower(self, prog, track):
        prog.add(_lib.MIYAMOTO, [self._G * self.m, self.a, self.b], track)


class NFWPotential(Potential):                        # potential.py:74-84
    def __init__(self, m, r_s, units=None):
        super().__init__(units, {'m': m, 'r_s': r_s})

    def _lower(self, prog, track):
        prog.add(_lib.NFW, [self._G * self.m, self.r_s], track)


class TriaxialNFWPotential(Potential):                # potential.py:86-97
    def __init__(self, m, r_s, q1=1.0, q2=1.0, q3=1.0, units=None):
        super().__init__(units, {'m': m, 'r_s': r_s, 'q1': q1, 'q2': q2, 'q3': q3})

    def _lower(self, prog, track):
        prog.add(_lib.TRIAXNFW, [self._G * self.m, self.r_s, self.q1, self.q2, self.q3], track)


class Isochrone(Potential):                           # potential.py:114-122
    def __init__(self, m, a, units=None):
        super().__init__(units, {'m': m, 'a': a})

    def _lower(self, prog, track):
        prog.add(_lib.ISOCHRONE, [self._G * self.m, self.a], track)


class PlummerPotential(Potential):                    # potential.py:124-130
    def __init__(self, m, r_s, units=None):
        super().__init__(units, {'m': m, 'r_s': r_s})

    def _lower(self, prog, track):
        prog.add(_lib.PLUMMER, [self._G * self.m, self.r_s], track)


class HernquistPotential(Potential):                  # potential.py:132-138
    def __init__(self, m, r_s, soft=0.0, units=None):
        super().__init__(units, {'m': m, 'r_s': r_s, 'soft': soft})

    def _lower(self, prog, track):
        prog.add(_lib.HERNQUIST, [self._G * self.m, self.r_s, self.soft], track)


class BarPotential(Potential):                        # potential.py:178-198
    """Long & Murali (1992) bar rotating about z with pattern speed Omega.  Derivatives on the device by Taylor jets (csrc/ssb_jet.cuh), as the
    reference obtains them by autodiff; the component runs on the interpreter path (not in a fused signature)."""

    def __init__(self, m, a, b, c, Omega, units=None):
        super().__init__(units, {'m': m, 'a': a, 'b': b, 'c': c, 'Omega': Omega})

    def _lower(self, prog, track):
        prog.add(_lib.BAR, [self._G * self.m, self.a, self.b, self.c, self.Omega], track)


class DehnenBarPotential(Potential):                  # potential.py:200-222
    def __init__(self, alpha, v0, R0, Rb, phib, Omega, units=None):
        super().__init__(units, {'alpha': alpha, 'v0': v0, 'R0': R0, 'Rb': Rb, 'phib': phib, 'Omega': Omega})

    def _lower(self, prog, track):
        prog.add(_lib.DEHNEN_BAR, [self.alpha, self.v0, self.R0, self.Rb, self.phib, self.Omega], track)


class MN3ExponentialDiskPotential(Potential):         # potential.py:224-280 (sum of three Miyamoto-Nagai disks)
    _K_pos_dens = np.array([
        [0.0036, -0.0330, 0.1117, -0.1335, 0.1749], [-0.0131, 0.1090, -0.3035, 0.2921, -5.7976],
        [-0.0048, 0.0454, -0.1425, 0.1012, 6.7120], [-0.0158, 0.0993, -0.2070, -0.7089, 0.6445],
        [-0.0319, 0.1514, -0.1279, -0.9325, 2.6836], [-0.0326, 0.1816, -0.2943, -0.6329, 2.3193]])
    _K_neg_dens = np.array([
        [-0.0090, 0.0640, -0.1653, 0.1164, 1.9487], [0.0173, -0.0903, 0.0877, 0.2029, -1.3077],
        [-0.0051, 0.0287, -0.0361, -0.0544, 0.2242], [-0.0358, 0.2610, -0.6987, -0.1193, 2.0074],
        [-0.0830, 0.4992, -0.7967, -1.2966, 4.4441], [-0.0247, 0.1718, -0.4124, -0.5944, 0.7333]])

    def __init__(self, m, h_R, h_z, units=None, positive_density=True, sech2_z=True):
        super().__init__(units, {'m': m, 'h_R': h_R, 'h_z': h_z})
        self.positive_density, self.sech2_z = positive_density, sech2_z
        K = self._K_pos_dens if positive_density else self._K_neg_dens
        hzR = h_z / h_R
        if sech2_z:
            b_hR = -0.033 * hzR ** 3 + 0.262 * hzR ** 2 + 0.659 * hzR
        else:
            b_hR = -0.269 * hzR ** 3 + 1.08 * hzR ** 2 + 1.092 * hzR
        param_vec = K @ np.array([b_hR ** 4, b_hR ** 3, b_hR ** 2, b_hR, 1.0])
        self._ms, self._as, self._b = param_vec[:3] * m, param_vec[3:] * h_R, b_hR * h_R

    def _lower(self, prog, track):
        for i in range(3):
            prog.add(_lib.MIYAMOTO, [self._G * self._ms[i], self._as[i], self._b], track)


class Potential_Combine(Potential):                   # potential.py:1279-1296
    def __init__(self, potential_list, units=None):
        super().__init__(units, {'potential_list': potential_list})

    def _lower(self, prog, track):
        for p in self.potential_list:
            p._lower(prog, track)


class GalaMilkyWayPotential(Potential):               # potential.py:390-418
    def __init__(self, units=None):
        super().__init__(units, {'params': None})
        self.m_disk, self.a_disk, self.b_disk = 6.80e10, 3.0, 0.28
        self.m_bulge, self.c_bulge = 5e9, 1.0
        self.m_nucleus, self.c_nucleus = 1.71e9, 0.07
        self.m_halo, self.r_s_halo = 5.4e11, 15.62
        pot_disk = MiyamotoNagaiDisk(m=self.m_disk, a=self.a_disk, b=self.b_disk, units=units)
        pot_bulge = HernquistPotential(m=self.m_bulge, r_s=self.c_bulge, units=units)
        pot_nucleus = HernquistPotential(m=self.m_nucleus, r_s=self.c_nucleus, units=units)
        pot_halo = NFWPotential(m=self.m_halo, r_s=self.r_s_halo, units=units)
        self.pot = Potential_Combine(potential_list=[pot_disk, pot_bulge, pot_nucleus, pot_halo], units=units)

    def _lower(self, prog, track):
        self.pot._lower(prog, track)


class TimeDepTranslatingPotential(Potential):         # potential.py:448-462
    """pot evaluated at xyz - center_spl(t).  center_spl must be a tabulated track (LinearTrack / CubicTrack or an
    interpax-like object with .x, .f, .method)."""

    def __init__(self, pot, center_spl, units=None):
        super().__init__(units, {'pot': pot, 'center_spl': center_spl})
        self._track = rt.as_track(center_spl)

    def _lower(self, prog, track):
        _no_nested(track)
        self.pot._lower(prog, prog.add_track(self._track))


class PerturberSetPotential(Potential):
    """N moving spheres of one profile (func = PlummerPotential / HernquistPotential / NFWPotential) whose centres are tabulated on ONE time
    grid: m[N], r_s[N] (or scalars), t[nk], centers[nk, N, 3]; linear interpolation in time (linear extrapolation outside, like the
    reference's RegularGridInterpolator tables, potential.py:581-600).  The same field as Potential_Combine of N
    TimeDepTranslatingPotential(func(m_i, r_s_i), LinearTrack(t, centers[:, i])) objects (potential.py:448-462) - which is also lowered to
    this form automatically when it would exceed the component / track limits of a program - without the limit of 12 components / 4 tracks:
    BASELINE config 5 (tracers in the field of 100 moving perturbers)."""

    def __init__(self, func, m, r_s, t, centers, units=None):
        super().__init__(units, {'func': func, 'm': m, 'r_s': r_s, 't': t, 'centers': centers})
        m = np.atleast_1d(np.asarray(m, dtype=np.float64))
        self._arrays = rt.PerturberArrays(_profile_of(func), self._G * m, np.broadcast_to(np.asarray(r_s, dtype=np.float64), m.shape), t, centers)

    def _lower(self, prog, track):
        _no_nested(track)
        if prog.growth:
            raise NotImplementedError("GrowingPotential around a perturber set is not supported")
        prog.add_perturbers(self._arrays)


class UniformAcceleration(Potential):                 # potential.py:480-502
    """Spatially uniform acceleration: gradient = d velocity_func / dt (slope of the tabulated velocity track)."""

    def __init__(self, velocity_func=None, units=None):
        super().__init__(units, {'velocity_func': velocity_func})
        self._track = rt.as_track(velocity_func)

    def potential(self, xyz, t):
        raise NotImplementedError

    def _lower(self, prog, track):
        prog.add(_lib.UNIFORM_ACC, [], prog.add_track(self._track))


class MW_LMC_Potential(Potential):                    # potential.py:555-662
    """MW (Hernquist bulge + Miyamoto-Nagai disk + NFW halo) + translating NFW LMC + uniform frame acceleration.

    The reference loads its 1000-knot tables from pickled jax arrays in its own package data
    (data/LMC_MW_potential/*.npy); those files are not part of this repository.  Pass the tables explicitly
    (`t_lmc, xyz_lmc, t_mw, vel_mw`) or `data_dir=` pointing at a streamsculptor checkout whose tables can be read.
    """

    def __init__(self, units=None, t_lmc=None, xyz_lmc=None, t_mw=None, vel_mw=None, data_dir=None):
        super().__init__(units, {'params': None})
        if t_lmc is None:
            t_lmc, xyz_lmc, t_mw, vel_mw = _load_mw_lmc_tables(data_dir)
        self.LMC_pos = LinearTrack(t_lmc, xyz_lmc)
        self.LMC_vel = LinearTrack(t_mw, vel_mw)
        pot_bulge = HernquistPotential(m=5e9, r_s=1.0, units=units)
        pot_disk = MiyamotoNagaiDisk(m=5.0e10, a=3.0, b=0.3, units=units)
        pot_halo = NFWPotential(m=5.4e11, r_s=15.62, units=units)
        self.pot_MW = Potential_Combine([pot_bulge, pot_disk, pot_halo], units=units)
        massLMC = .85e11
        radiusLMC = (massLMC / 1e11) ** 0.6 * 8.5
        self.pot_LMC = NFWPotential(m=massLMC, r_s=radiusLMC, units=units)
        self.translating_LMC_pot = TimeDepTranslatingPotential(pot=self.pot_LMC, center_spl=self.LMC_pos, units=units)
        self.unif_acc = UniformAcceleration(velocity_func=self.LMC_vel, units=units)
        self.total_pot = Potential_Combine([self.pot_MW, self.translating_LMC_pot, self.unif_acc], units=units)

    def LMC_center_spline(self, t):
        return self.LMC_pos(t)

    def MW_velocity_func(self, t):
        return self.LMC_vel(t)

    def potential(self, xyz, t):
        raise NotImplementedError("Potential not implemented, force is non-conservative")

    def _lower(self, prog, track):
        self.total_pot._lower(prog, track)


def _load_mw_lmc_tables(data_dir):
    """Read the reference's pickled tables (data/LMC_MW_potential/*.npy hold dicts of jax Arrays) without importing jax:
    a jax Array pickles as _reconstruct_array(numpy_reconstruct, args, ndarray_state, aval_state), which a stub rebuilds."""
    import pickle
    if data_dir is None:
        raise FileNotFoundError("MW_LMC_Potential needs its motion tables: pass t_lmc/xyz_lmc/t_mw/vel_mw, or data_dir= pointing at "
                                "streamsculptor/data/LMC_MW_potential of a reference checkout")

    def _rebuild(fun, args, arr_state, *rest):
        arr = fun(*args)
        arr.__setstate__(arr_state)
        return arr

    class _Stub(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] in ("jax", "jaxlib"):
                return _rebuild
            return super().find_class(module, name)

    def load(fn):
        with open(os.path.join(data_dir, fn), "rb") as f:
            major, minor = np.lib.format.read_magic(f)
            (np.lib.format.read_array_header_1_0 if (major, minor) == (1, 0) else np.lib.format.read_array_header_2_0)(f)
            obj = _Stub(f).load()
            return obj.item() if isinstance(obj, np.ndarray) else obj     # np.save(dict) stores a 0-d object array
    lmc, mw = load("LMC_motion_dict.npy"), load("MW_motion_dict.npy")
    return (np.asarray(lmc['flip_tsave'], dtype=np.float64), np.asarray(lmc['flip_trajLMC'], dtype=np.float64)[:, :3],
            np.asarray(mw['flip_tsave'], dtype=np.float64), np.asarray(mw['flip_traj'], dtype=np.float64)[:, 3:6])


# ---- subhalo ensembles (potential.py:802-956, 1110-1268) ------------------------------------------------------------
def _profile_of(func):
    name = func if isinstance(func, str) else getattr(func, "__name__", type(func).__name__)
    table = {"PlummerPotential": _lib.PROFILE_PLUMMER, "HernquistPotential": _lib.PROFILE_HERNQUIST, "NFWPotential": _lib.PROFILE_NFW}
    if name not in table:
        raise NotImplementedError(f"subhalo profile {name!r} has no closed-form device implementation (Plummer, Hernquist, NFW only)")
    return table[name]


class _SubhaloLineBase(Potential):
    _dradius = False

    def _setup(self, profile, m, r_s):
        self._arrays = rt.SubhaloArrays(profile, self._G, m, r_s, self.subhalo_x0, self.subhalo_v, self.subhalo_t0, self.t_window)

    def _lower(self, prog, track):
        if self._dradius:
            raise NotImplementedError("d/dr_s subhalo potentials only enter through the perturbation field (fields.py:200)")
        prog.add_subhalos(self._arrays, track)

    def potential_per_SH(self, xyz, t):               # potential.py:832-850 etc.
        phi, _ = rt.subhalo_eval(self._arrays, self._dradius, xyz, t)
        return phi.cpu().numpy()

    def gradient_per_SH(self, xyz, t):                # jacfwd(potential_per_SH) (perturbative.py:40-41, 695-696)
        _, g = rt.subhalo_eval(self._arrays, self._dradius, xyz, t)
        return g.cpu().numpy()

    def potential(self, xyz, t):
        if self._dradius:
            return float(self.potential_per_SH(xyz, t).sum())
        return super().potential(xyz, t)

    def gradient(self, xyz, t):
        if self._dradius:
            return self.gradient_per_SH(xyz, t).sum(axis=0)
        return super().gradient(xyz, t)


class SubhaloLinePotential(_SubhaloLineBase):         # potential.py:802-850 (Plummer spheres)
    def __init__(self, m, a, subhalo_x0, subhalo_v, subhalo_t0, t_window, units=None):
        super().__init__(units, {'m': m, 'a': a, 'subhalo_x0': subhalo_x0, 'subhalo_v': subhalo_v, 'subhalo_t0': subhalo_t0, 't_window': t_window})
        self._setup(_lib.PROFILE_PLUMMER, m, a)


class SubhaloLinePotential_dRadius(_SubhaloLineBase):  # potential.py:852-904
    _dradius = True

    def __init__(self, m, a, subhalo_x0, subhalo_v, subhalo_t0, t_window, units=None):
        super().__init__(units, {'m': m, 'a': a, 'subhalo_x0': subhalo_x0, 'subhalo_v': subhalo_v, 'subhalo_t0': subhalo_t0, 't_window': t_window})
        self._setup(_lib.PROFILE_PLUMMER, m, a)


class SubhaloLinePotentialCustom_fromFunc(_SubhaloLineBase):   # potential.py:1161-1213
    def __init__(self, func, m, r_s, subhalo_x0, subhalo_v, subhalo_t0, t_window, units=None):
        super().__init__(units, {'func': func, 'm': m, 'r_s': r_s, 'subhalo_x0': subhalo_x0, 'subhalo_v': subhalo_v,
                                 'subhalo_t0': subhalo_t0, 't_window': t_window})
        self._setup(_profile_of(func), m, r_s)


class SubhaloLinePotentialCustom_dRadius_fromFunc(_SubhaloLineBase):   # potential.py:1215-1268
    _dradius = True

    def __init__(self, func, m, r_s, subhalo_x0, subhalo_v, subhalo_t0, t_window, units=None):
        super().__init__(units, {'func': func, 'm': m, 'r_s': r_s, 'subhalo_x0': subhalo_x0, 'subhalo_v': subhalo_v,
                                 'subhalo_t0': subhalo_t0, 't_window': t_window})
        self._setup(_profile_of(func), m, r_s)


class SubhaloLinePotential_Custom(_SubhaloLineBase):   # potential.py:908-956: ONE potential object `pot` shared by every subhalo
    """`pot` must be a Plummer / Hernquist / NFW instance (its m and r_s are used for all subhalos)."""

    def __init__(self, pot, subhalo_x0, subhalo_v, subhalo_t0, t_window, units=None):
        super().__init__(units, {'pot': pot, 'subhalo_x0': subhalo_x0, 'subhalo_v': subhalo_v, 'subhalo_t0': subhalo_t0, 't_window': t_window})
        n = len(np.atleast_1d(np.asarray(subhalo_t0)))
        if float(getattr(pot, 'soft', 0.0) or 0.0) != 0.0:
            # the reference evaluates self.pot.potential(), softening included (potential.py:136-138, 908-956); the subhalo arrays of the
            # kernels (ssb_subhalos) carry no softening, so a softened profile must not be evaluated silently with soft = 0
            raise NotImplementedError("SubhaloLinePotential_Custom: a softened `pot` (soft != 0) is not supported by the subhalo kernels; "
                                      "use soft=0 or add the subhalo as a translating HernquistPotential component")
        rs = getattr(pot, 'r_s', None)
        if rs is None:
            rs = getattr(pot, 'a')
        self._arrays = rt.SubhaloArrays(_profile_of(type(pot)), pot._G, np.full(n, float(pot.m)), np.full(n, float(rs)), subhalo_x0, subhalo_v,
                                        subhalo_t0, t_window)


def _track_of_interp(interp_func, n_knots=4097):
    """`interp_func` in the reference is a diffrax dense Solution (`.evaluate(t)[:3]`, potential.py:152).  The kernels need a tabulated
    track: LinearTrack / CubicTrack / interpax-like objects pass through; a dense Solution of this package is re-sampled on `n_knots`
    uniform knots into a cubic-Hermite track."""
    try:
        return rt.as_track(interp_func)
    except NotImplementedError:
        pass
    dense = getattr(interp_func, "_dense", None)
    if dense is None:
        raise NotImplementedError("interp_func must be a tabulated track or a dense Solution returned by integrate_orbit(dense=True)")
    tk = np.linspace(min(dense.t0, dense.t1), max(dense.t0, dense.t1), n_knots)
    return CubicTrack(tk, np.asarray(interp_func.evaluate(tk))[:, :3])


class ProgenitorPotential(Potential):                 # potential.py:140-153
    """prog_pot(m, r_s) centred on the progenitor's interpolated trajectory."""

    def __init__(self, m, r_s, interp_func, prog_pot, units=None):
        super().__init__(units, {'m': m, 'r_s': r_s, 'interp_func': interp_func, 'prog_pot': prog_pot})
        self.prog_pot = prog_pot(m=self.m, r_s=self.r_s, units=units)
        self._track = _track_of_interp(interp_func)

    def _lower(self, prog, track):
        _no_nested(track)
        self.prog_pot._lower(prog, prog.add_track(self._track))


# ---- reference classes deliberately outside the B200 hot path -----------------------------------------------------
def _out_of_scope(name, why):
    class _X(Potential):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: {why} (outside the B200 hot path; see DESIGN.md 'Out of scope')")
    _X.__name__ = name
    return _X


CustomPotential = _out_of_scope("CustomPotential", "arbitrary Python potential functions cannot run inside a CUDA kernel")
ZhaoPotential = _out_of_scope("ZhaoPotential", "needs incomplete beta functions")
PowerLawCutoffPotential = _out_of_scope("PowerLawCutoffPotential", "needs incomplete gamma functions")
BovyMWPotential2014 = _out_of_scope("BovyMWPotential2014", "contains PowerLawCutoffPotential")
class GrowingPotential(Potential):                    # potential.py:464-477
    """pot.potential(xyz, t) * growth_func(t).  growth_func must be TABULATED: a LinearTrack / CubicTrack whose first column is the growth
    factor, an interpax-like object with .x / .f (1-D), or a pair (t, factor) of arrays (linear interpolation).  The wrapped potential may be
    any sum of the analytic components (and translating ones); force-only and subhalo-ensemble components are not supported inside it."""

    def __init__(self, pot, growth_func, units=None):
        super().__init__(units, {'pot': pot, 'growth_func': growth_func})
        self._gtrack = _scalar_track(growth_func)

    def _lower(self, prog, track):
        if prog.growth:
            raise NotImplementedError("nested GrowingPotential is not implemented")
        prog.growth = prog.add_track(self._gtrack) + 1
        try:
            self.pot._lower(prog, track)
        finally:
            prog.growth = 0


def _scalar_track(f):
    """Tabulated scalar function of time -> a Track whose first column carries it."""
    if isinstance(f, rt.Track):
        return f
    if isinstance(f, (tuple, list)) and len(f) == 2:
        t, y = np.asarray(f[0], dtype=np.float64), np.asarray(f[1], dtype=np.float64)
        return rt.Track(_lib.TRACK_LINEAR, t, np.stack([y, np.zeros_like(y), np.zeros_like(y)], axis=1))
    if hasattr(f, "x") and hasattr(f, "f"):
        y = np.asarray(f.f, dtype=np.float64).reshape(len(np.asarray(f.x)), -1)[:, 0]
        kind = _lib.TRACK_LINEAR if getattr(f, "method", "cubic") == "linear" else _lib.TRACK_CUBIC
        return rt.Track(kind, np.asarray(f.x), np.stack([y, np.zeros_like(y), np.zeros_like(y)], axis=1))
    raise NotImplementedError("growth_func must be tabulated (LinearTrack / CubicTrack, an interpax-like object, or a (t, factor) pair); "
                              "arbitrary Python callables cannot run inside the CUDA kernels")
