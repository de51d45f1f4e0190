"""Host-side mirror of the reference's ``init_`` module (parameter source of the hot path).

Same public names and values as /root/reference/init_.py so that code written against
the reference (``init_.Parameters``, ``init_.GaAs``, ``init_.Device``, ``init_.init_*``)
keeps working; the heavy lifting (10^6..10^9-particle ensembles) happens on the GPU in
:mod:`monte_carlo_carrier_transport_b200.engine`.

Reference map: Layer init_.py:15-32, where_am_i :36-56, Parameters :58-70, GaAs :73-105,
Device :108-157, init_energy_df :160-166, init_elec_field :168-173, init_mesh :175-187,
init_rho :189-191, init_coords :194-234, ``laser``/``init_photoex`` :236-280 (host mirrors;
the device-side injection is ``engine.Ensemble.inject_photoex``).
"""
from __future__ import annotations

import itertools

import numpy as np


class Layer:
    """One slab of the layer stack along z (init_.py:15-32). Lengths in metres, doping in cm^-3."""

    def __init__(self, matl, dope, lx, ly, lz):
        self.matl, self.dope, self.lx, self.ly, self.lz = matl, dope, lx, ly, lz


def where_am_i(layers, dist):
    """z measured from the start of layer 0 -> ``{'current_layer': index, 'dist': depth}``.

    Same contract and error behaviour as init_.py:36-56 / main_.py:156-177: a point
    exactly on an interface belongs to the lower layer; outside the stack raises
    ``ValueError``."""
    if dist < 0:
        raise ValueError("Point is outside all layers")
    depth = dist
    for index, layer in enumerate(layers):
        if depth <= layer.lz:
            return {"current_layer": index, "dist": depth}
        depth -= layer.lz
    raise ValueError("Point is outside all layers. Distance = " + str(depth))


class Parameters:
    """Simulation parameters and (rounded) physical constants, init_.py:58-70."""

    dt = 2E-15
    pts = 50
    inf = float("inf")
    kT = 0.02585
    m_0 = 9.11E-31
    q = 1.602E-19
    hbar = 1.055E-34
    hbar_eVs = 6.582E-16
    eps0 = 8.854E-12


class GaAs:
    """Three-valley GaAs, init_.py:73-105. Valley keys: 0 = Gamma, 1 = L, 2 = X."""

    EG = 1.424
    eps_0 = 13.1 * Parameters.eps0
    eps_inf = 10.9 * Parameters.eps0
    rho = 5360
    E_phonon = 0.03536
    vs = 5.22E3
    dope = 1E16
    alpha = 1.1E6
    tot_scat = 4E14
    mass = {0: 0.063 * Parameters.m_0, 1: 0.170 * Parameters.m_0, 2: 0.58 * Parameters.m_0}
    nonparab = {0: 0.610, 1: 0.461, 2: 0.204}
    DA = {0: 7.01, 1: 9.2, 2: 9.0}
    E_val = {"GL": 0.29, "LG": 0.29, "GX": 0.48, "XG": 0.48, "LX": 0.19, "XL": 0.19}
    B = {0: 1, 1: 4, 2: 3}
    defpot = {"GL": 1.8E8, "LG": 1.8E8, "GX": 10E8, "XG": 10E8, "LX": 1E8, "XL": 1E8,
              "LL": 1E8, "XX": 10E8}
    EP = {"GL": 0.0278, "LG": 0.0278, "GX": 0.0299, "XG": 0.0299, "LX": 0.0293, "XL": 0.0293,
          "LL": 0.029, "XX": 0.0299}


class Device:
    """Geometry / mesh derivation, init_.py:108-157.

    The reference evaluates this in the class body at import time; here the same
    quantities are (re)derived by :meth:`configure`, which is called once at import
    with the reference's shipped device so that the class attributes match."""

    @staticmethod
    def inv_debye_len(layer):
        """init_.py:110-123"""
        return np.sqrt(Parameters.q * layer.dope * 1E2 ** 3 / (layer.matl.eps_0 * Parameters.kT))

    @classmethod
    def configure(cls, layers=None, num_carr=10000, elec_field=(0, 0, -5000), mean_energy=0.1):
        if layers is None:
            layers = [Layer(matl=GaAs, dope=1E16, lx=3E-6, ly=4.326178398780776e-08, lz=900E-9)]
        cls.layers = list(layers)
        cls.layer0 = cls.layers[0]
        cls.Materials = [layer.matl for layer in cls.layers]
        cls.num_carr = num_carr
        cls.elec_field = np.array(elec_field)
        cls.dim = [np.array([layer.lx, layer.ly, layer.lz]) for layer in cls.layers]
        # cell size ~ Debye length (init_.py:135-143)
        seg0 = np.array([(d * cls.inv_debye_len(layer)).astype(int)
                         for d, layer in zip(cls.dim, cls.layers)])
        cls.dl = np.max([d / s for d, s in zip(cls.dim, seg0)], axis=0)
        cls.seg = (cls.dim / cls.dl).astype(int)
        cls.mean_energy = mean_energy
        cls.tot_dim = np.append(np.max(cls.dim, axis=0)[:2], np.sum(cls.dim, axis=0)[2])
        cls.tot_seg = np.append(np.max(cls.seg, axis=0)[:2], np.sum(cls.seg, axis=0)[2]).astype(int)
        cls.vol = np.prod(cls.dim, axis=1)
        # super-particle weight per layer; doping is in cm^-3 (init_.py:153-157)
        cls.charge = np.array([(layer.dope * v * (1e2) ** 3) / num_carr
                               for layer, v in zip(cls.layers, cls.vol)])
        return cls


Device.configure()


def init_energy_df(max_nrg=2, div=10000):
    """Energy discretisation as a one-column DataFrame ``Ek`` (init_.py:160-166)."""
    import pandas as pd

    return pd.DataFrame(np.linspace(0, max_nrg, div), columns=["Ek"])


def init_elec_field():
    """Constant field tiled over the mesh, shape (nx, ny, nz, 3) (init_.py:168-173)."""
    return np.tile(Device.elec_field, np.append(Device.tot_seg.astype(int), 1))


def _cell_centres():
    return [np.arange(1, Device.tot_seg[j] * 2 + 1, 2) * Device.dl[j] / 2 for j in range(3)]


def init_mesh():
    """Cell-centre coordinates, z fastest (init_.py:175-187)."""
    import pandas as pd

    return pd.DataFrame(list(itertools.product(*_cell_centres())))


def init_rho(dfmesh, dope, dl):
    """Background (donor) charge count per cell, negative (init_.py:189-191)."""
    return np.ones(len(dfmesh)) * dope * np.prod(dl) * (-100 ** 3)


def init_coords(layers=None, num_carr=None, mean_nrg=None):
    """Initial ensemble ``(N, 6)`` = kx, ky, kz, x, y, z on the HOST with the reference's
    global-stream draw order (init_.py:194-234): x, y, z uniform; energies
    ``maxwell(loc=mean*kT, scale=1e-7)``; then per particle two ``Normal(0, 2*pi)`` angles.
    For large ensembles use ``engine.Ensemble.init_ensemble`` (device-side Philox)."""
    from scipy.stats import maxwell

    layers = Device.layers if layers is None else layers
    num_carr = Device.num_carr if num_carr is None else num_carr
    mean_nrg = Device.mean_energy if mean_nrg is None else mean_nrg
    n = int(num_carr)
    box = np.max(Device.dim, axis=0)
    pos = [np.random.uniform(0, box[a], n) for a in range(3)]
    e = maxwell.rvs(loc=mean_nrg * Parameters.kT, scale=1E-7, size=num_carr)
    coords = np.zeros((n, 6))
    for a in range(3):
        coords[:, 3 + a] = pos[a]
    for i in range(n):
        mat = layers[where_am_i(layers, pos[2][i])["current_layer"]].matl
        k = np.sqrt(2 * mat.mass[0] * Parameters.q * e[i] / Parameters.hbar ** 2)
        alpha = np.random.normal(0, 2 * np.pi)
        beta = np.random.normal(0, 2 * np.pi)
        coords[i, 0] = k * np.cos(alpha) * np.sin(beta)
        coords[i, 1] = k * np.sin(alpha) * np.sin(beta)
        coords[i, 2] = k * np.cos(beta)
    return coords


class laser:
    """Excitation pulse parameters, init_.py:236-242.  The accelerated injection is
    ``engine.Ensemble.inject_photoex`` / ``main_.run(photoex=True)``."""

    laser_ex = 800
    laser_pow = 0.1E-3
    laser_std = 0.5E-6
    laser_t = 100E-15
    laser_eff = (laser_pow * laser_t / (Parameters.q * (1240 / laser_ex - GaAs.EG))) / 10000
    t0 = 1.5E-12


def init_photoex(layers, num_carr, nrg=1240 / laser.laser_ex - GaAs.EG):
    """Photo-excited carriers ``(M, 6)``, M <= num_carr (init_.py:244-280): exponential depth
    profile (carriers drawn beyond the slab are dropped), Gaussian x and y, mono-energetic,
    ``Normal(0, 2*pi)`` angles; same np.random draw order as the reference."""
    n = int(num_carr)
    z = np.random.exponential(min(1 / GaAs.alpha, Device.tot_dim[2]), n)
    keep = z < Device.tot_dim[2]
    z = z[keep]
    x = np.random.normal(0, laser.laser_std, n)[keep]
    y = np.random.normal(0, laser.laser_std, n)[keep]
    coords = np.zeros((len(z), 6))
    for i in range(len(z)):
        mat = layers[where_am_i(layers, z[i])["current_layer"]].matl
        k = np.sqrt(2 * mat.mass[0] * Parameters.q * nrg / Parameters.hbar ** 2)
        alpha = np.random.normal(0, 2 * np.pi)
        beta = np.random.normal(0, 2 * np.pi)
        coords[i] = [k * np.cos(alpha) * np.sin(beta), k * np.sin(alpha) * np.sin(beta), k * np.cos(beta),
                     x[i], y[i], z[i]]
    return coords
