"""ImpactGenerator of the reference (GenerateImpactParams.py:9-215; formalism of Yoon+2011, Erkal+2016): subhalo impact parameters for the
production driver (generate_derivs.py:153-204).

What is heavy runs on the device: the window means of the stream around every sampled phi1 (one sort + prefix sums in torch instead of
NumImpacts masked passes over the stream, GenerateImpactParams.py:151-163) and the NumImpacts backward orbit solves of the stream patches
to their impact times (ONE batched K1 launch instead of a vmap of integrate_orbit, GenerateImpactParams.py:191).  The random numbers follow
the jax.random recipes of the reference call by call (threefry split / uniform / normal / choice on the host: O(NumImpacts) integers);
split and randint are pinned by the oracle's recipes (golden D8 hangs on them), uniform by sample_from_1D_pdf's checks; normal and choice
have no reference output here and are covered by distribution tests.
"""
import numpy as np

from . import _runtime as rt
from .solvers import Dopri8
from .streamhelpers import _threefry2x32, compute_length_oscillations, compute_stream_length, jax_uniform, key_words, sample_from_1D_pdf

SIGMA_180_KMS = 0.18408818970766738          # (180 km/s).to(kpc/Myr)  (GenerateImpactParams.py:24)


# ---- jax.random on the host (threefry2x32; streamhelpers has the generator and uniform) ---------------------------------------------
def jax_split(key, n):
    """jax.random.split(key, n): [n, 2] key words."""
    k0, k1 = key_words(key)
    c = np.arange(n, dtype=np.uint64)
    o0, o1 = _threefry2x32(k0, k1, c, c + np.uint64(n))
    flat = np.concatenate([o0, o1]).astype(np.uint32)
    return flat.reshape(n, 2)


def jax_randint(key, n, minval, maxval):
    """jax.random.randint(key, (n,), minval, maxval) with 64-bit integers (x64 enabled, as the reference runs): split the key, 64 random
    bits from each half, multiply-mod combination.  Pinned by the oracle's recipe (golden D8 depends on it, tests/test_host_cpu.py)."""
    k1, k2 = jax_split(key, 2)
    j = np.arange(n, dtype=np.uint64)

    def bits64(k):
        hi, lo = _threefry2x32(int(k[0]), int(k[1]), j, j + np.uint64(n))
        return [(int(h) << 32) | int(l) for h, l in zip(hi, lo)]
    span = max(int(maxval) - int(minval), 1)
    mult = ((1 << 32) % span) ** 2 % span
    return np.array([int(minval) + ((a % span) * mult + (b % span)) % span for a, b in zip(bits64(k1), bits64(k2))], dtype=np.int64)


def jax_uniform_range(key, n, minval, maxval):
    """jax.random.uniform(key, (n,), minval, maxval) in float64."""
    minval, maxval = np.broadcast_to(np.asarray(minval, dtype=np.float64), (n,)), np.broadcast_to(np.asarray(maxval, dtype=np.float64), (n,))
    return np.maximum(minval, jax_uniform(key, n) * (maxval - minval) + minval)


def jax_normal(key, n):
    """jax.random.normal(key, (n,)) in float64: sqrt(2) erfinv(u), u uniform in (nextafter(-1, 0), 1)."""
    from scipy.special import erfinv
    lo = np.nextafter(-1.0, 0.0)
    return np.sqrt(2.0) * erfinv(jax_uniform_range(key, n, lo, 1.0))


def jax_choice(key, a, n, p):
    """jax.random.choice(key, a, (n,), replace=True, p=p): inverse CDF with r = p_cuml[-1] (1 - uniform)."""
    cum = np.cumsum(np.asarray(p, dtype=np.float64))
    r = cum[-1] * (1.0 - jax_uniform(key, n))
    return np.asarray(a)[np.minimum(np.searchsorted(cum, r), len(cum) - 1)]


class ImpactGenerator:
    def __init__(self, pot, tobs, stream, stream_phi1, stripping_times, prog_today, phi1window=0.1, NumImpacts=1, tImpactBounds=None,
                 bImpact_bounds=(0, 1.0), sigma=SIGMA_180_KMS, phi1_bounds=None, phi1_exclude=(1.0, 1.0), stream_length=None, seednum=0, shared=None):
        """shared: a dict the caller keeps between generators built on the SAME stream (the production driver builds one per batch): the
        stream's device copy, its sorted prefix sums and the length-oscillation table are computed by the first one and reused."""
        self.pot = pot
        self.tobs = float(tobs)
        self.NumImpacts = int(NumImpacts)
        self.phi1window, self.sigma, self.seednum = float(phi1window), float(sigma), int(seednum)
        self.prog_today = np.asarray(prog_today, dtype=np.float64)
        self.phi1_exclude = [float(phi1_exclude[0]), float(phi1_exclude[1])]
        self.keys = jax_split(self.seednum, 7)                                                     # GenerateImpactParams.py:44
        tt = rt.torch()
        cache = shared if shared is not None else {}
        if "stream" not in cache:
            cache["stream"] = (rt.to_dev(stream).reshape(-1, 6), rt.to_dev(stream_phi1).reshape(-1), rt.to_dev(stripping_times).reshape(-1))
        self.stream, self.stream_phi1, self.stripping_times = cache["stream"]
        phi1_h = np.asarray(stream_phi1, dtype=np.float64).reshape(-1) if not rt.is_dev(stream_phi1) else self.stream_phi1.cpu().numpy()
        self.tImpactBounds = [float(np.min(np.asarray(stripping_times.cpu() if rt.is_dev(stripping_times) else stripping_times))), 0.0] if tImpactBounds is None else [float(tImpactBounds[0]), float(tImpactBounds[1])]
        self.phi1_bounds = [float(phi1_h.min()), float(phi1_h.max())] if phi1_bounds is None else [float(phi1_bounds[0]), float(phi1_bounds[1])]
        if len(bImpact_bounds) == 2 and np.ndim(bImpact_bounds[0]) == 0 and np.ndim(bImpact_bounds[1]) == 0:
            self.b_low, self.b_high = float(bImpact_bounds[0]), float(bImpact_bounds[1])
        elif len(bImpact_bounds) == 2:                       # [low, high-array] as generate_derivs.py:174 passes it
            self.b_low = np.broadcast_to(np.asarray(bImpact_bounds[0], dtype=np.float64), (self.NumImpacts,))
            self.b_high = np.broadcast_to(np.asarray(bImpact_bounds[1], dtype=np.float64), (self.NumImpacts,))
        else:                                                # N_impacts x 2
            bb = np.asarray(bImpact_bounds, dtype=np.float64)
            self.b_low, self.b_high = bb[:, 0], bb[:, 1]
        osc_key = ("length_osc", abs(min(self.tImpactBounds)), None if stream_length is None else float(stream_length))
        if osc_key not in cache:
            stream_h = np.asarray(stream, dtype=np.float64).reshape(-1, 6) if not rt.is_dev(stream) else self.stream.cpu().numpy()
            length = compute_stream_length(stream=stream_h, phi1=phi1_h) if stream_length is None else stream_length    # GenerateImpactParams.py:64-68
            ind_break = len(self.stripping_times) // 2
            cache[osc_key] = compute_length_oscillations(pot=self.pot, prog_today=self.prog_today, first_stripped_lead=stream_h[0],
                                                         first_stripped_trail=stream_h[ind_break], t_age=abs(min(self.tImpactBounds)),
                                                         length_today=length)                     # GenerateImpactParams.py:74-83
        self.length_osc = cache[osc_key]
        if "sorted" not in cache:                                # sorted copy + prefix sums: window means in O(log N) per sample
            order = tt.argsort(self.stream_phi1)
            zero = tt.zeros((1, 7), dtype=tt.float64, device=self.stream.device)
            cache["sorted"] = (self.stream_phi1[order].contiguous(),
                               tt.cat([zero, tt.cumsum(tt.cat([self.stream[order], self.stripping_times[order, None]], dim=1), dim=0)]))
        self._phi1_sorted, self._cum = cache["sorted"]

    # GenerateImpactParams.py:87-105
    def w_parallel_sample(self, vs):
        return jax_normal(self.keys[0], self.NumImpacts) * self.sigma - vs

    def w_perpendicular_sample(self):
        prefac = np.sqrt(2 / np.pi) / self.sigma ** 3
        w = np.linspace(-7 * self.sigma, 7 * self.sigma, 10_000)
        prob = prefac * w ** 2 * np.exp(-w ** 2 / (2 * self.sigma ** 2))
        return jax_choice(self.keys[1], w, self.NumImpacts, prob / prob.sum())

    def sample_impact_params(self):                           # GenerateImpactParams.py:110-147
        keys = jax_split(self.keys[-1], 4)
        n = self.NumImpacts
        bImpact = jax_uniform_range(keys[0], n, self.b_low, self.b_high)
        tImpact = sample_from_1D_pdf(x=self.length_osc['ts'], y=self.length_osc['length_func'], key=keys[1], num_samples=n)
        if abs(self.phi1_exclude[1] - self.phi1_exclude[0]) > 0:
            seg1 = np.linspace(self.phi1_bounds[0], self.phi1_exclude[0], 1000)
            seg2 = np.linspace(self.phi1_exclude[1], self.phi1_bounds[1], 1000)
            l1, l2 = self.phi1_exclude[0] - self.phi1_bounds[0], self.phi1_bounds[1] - self.phi1_exclude[1]
            prob = np.hstack([np.full(1000, (l1 / (l1 + l2)) / 1000.0), np.full(1000, (l2 / (l1 + l2)) / 1000.0)])
            phi1 = jax_choice(keys[2], np.hstack([seg1, seg2]), n, prob)
        else:
            phi1 = jax_uniform_range(keys[2], n, self.phi1_bounds[0], self.phi1_bounds[1])
        perp = jax_uniform_range(keys[3], n, 0.0, 2 * np.pi)
        return {"bImpact": bImpact, "tImpact": tImpact, "phi1_samples": phi1, "perp_angle": perp}

    def get_particle_mean(self, phi1_0):
        """Mean phase-space position and mean stripping time of the stream particles with |phi1 - phi1_0| < phi1window (strict), for an array of
        phi1_0 (GenerateImpactParams.py:151-163): device tensors [n, 6], [n]; NaN rows where the window is empty, as the reference's 0/0."""
        tt = rt.torch()
        c = rt.to_dev(np.atleast_1d(np.asarray(phi1_0, dtype=np.float64)))
        lo = tt.searchsorted(self._phi1_sorted, c - self.phi1window, right=True)
        hi = tt.searchsorted(self._phi1_sorted, c + self.phi1window, right=False)
        cnt = (hi - lo).clamp(min=0).to(tt.float64)
        s = (self._cum[hi] - self._cum[lo]) / cnt[:, None]
        return s[:, :6].contiguous(), s[:, 6].contiguous()

    def get_subhalo_ImpactParams(self):                       # GenerateImpactParams.py:166-213
        tt = rt.torch()
        par = self.sample_impact_params()
        n = self.NumImpacts
        means, _tstrip = self.get_particle_mean(par["phi1_samples"])
        t_imp = rt.to_dev(par["tImpact"])
        # the stream patch of every impact integrated from tobs back to its impact time: one batched launch (Dopri8, 1e-7, dtmin 0.1)
        ctrl = rt.make_ctrl(Dopri8(), 1e-7, 1e-7, 0.1, None, 10_000)
        t0 = tt.full((n,), self.tobs, dtype=tt.float64, device=means.device)
        ys, status, _ = rt.orbit_integrate(self.pot, means, t0, t_imp, t_imp.reshape(-1, 1), ctrl, ts_per_orbit=1)
        W0 = ys[:, 0]
        x, v = W0[:, :3], W0[:, 3:]
        vs = tt.linalg.norm(v, dim=1)
        T = v / vs[:, None]
        B = tt.linalg.cross(x, v)
        B = B / tt.linalg.norm(B, dim=1)[:, None]
        N = tt.linalg.cross(B, T)
        # w_parallel_sample(v_s)[subidx] + v_s = normal[subidx] * sigma  (GenerateImpactParams.py:198)
        w_par = rt.to_dev(jax_normal(self.keys[0], n) * self.sigma)
        w_perp = rt.to_dev(self.w_perpendicular_sample())
        ang, b = rt.to_dev(par["perp_angle"]), rt.to_dev(par["bImpact"])
        V = w_par[:, None] * T + w_perp[:, None] * (-N * tt.sin(ang)[:, None] + B * tt.cos(ang)[:, None])
        b_hat = N * tt.cos(ang)[:, None] + B * tt.sin(ang)[:, None]
        X = x + b[:, None] * b_hat
        # the reference returns the first two; StreamPatch (the patch's state at its impact time) and status are extras
        return {"CartesianImpactParams": tt.cat([X, V], dim=1).cpu().numpy(), "ImpactFrameParams": par, "StreamPatch": W0.cpu().numpy(),
                "status": status.cpu().numpy()}
