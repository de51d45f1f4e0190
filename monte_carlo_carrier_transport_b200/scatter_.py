"""Host-side mirror of the reference's ``scatter_`` module.

What lives here (host, run once): the 27 scattering-rate formulas, ``scatter_mech`` and
``calc_scatter_table`` -- the inputs of the GPU hot path.
What does NOT live here: the per-event work.  ``calc_energy`` and ``scatter`` keep the
reference's signatures but execute on the GPU through the C-ABI (``emc_calc_energy``,
``emc_scatter``); there is no CPU fallback.

The rate formulas are generated from two descriptor tables instead of 27 hand-written
functions, but every expression keeps the reference's operation ORDER, so the tables
are bit-identical to ``scatter_.calc_scatter_table`` (tests/test_oracle_golden.py and
tests/test_host_mirrors.py pin the SHA-256 recorded from the unmodified reference).  Reference map: Bose_Einstein
scatter_.py:57-70, density_of_states :73-89, plasmon_freq :91-106, ADP :110-130,
POP_abs/ems :133-182, IV_* :185-580, IMP :583-624, cc :626-652, total/self :654-792,
angle samplers :795-889 (device code: csrc/emc_device.cuh ``sample_polar``), scatter_mech :892-951,
calc_scatter_table :954-984, scatter :986-1099.
"""
from __future__ import annotations

import numpy as np

from .init_ import Device, Parameters

Materials = Device.Materials

_VALLEY_LETTER = "GLX"

# sampler ids shared with include/emc.h (EMC_SAMPLER_*)
SAMPLER_SELF, SAMPLER_ISO, SAMPLER_POP_ABS, SAMPLER_POP_EMS, SAMPLER_ION = range(5)


def _materials():
    return Device.Materials


def Bose_Einstein(Ek):
    """scatter_.py:57-70"""
    return 1 / (np.exp(Ek / Parameters.kT) - 1)


def density_of_states(Ek, m):
    """scatter_.py:73-89"""
    return 2 * (2 * m) ** (1.5) * np.sqrt(Ek) / (4 * np.pi ** 2 * Parameters.hbar_eVs ** 3)


def plasmon_freq(dope, mat_i, val):
    """scatter_.py:91-106"""
    mat = _materials()[mat_i]
    return np.sqrt(Parameters.q ** 2 * dope * 1E2 ** 3 / (mat.eps_0 * mat.mass[val]))


def _wavevector(Ek, mat, val):
    return np.sqrt(2 * Parameters.q * Ek * mat.mass[val]) / Parameters.hbar


def _as_energy(Ek):
    return np.asarray(Ek, dtype=np.float64)


# ---------------------------------------------------------------------------------
# intravalley mechanisms
# ---------------------------------------------------------------------------------
def ADP(Ek, mat_i, val, T):
    """Acoustic deformation potential, scatter_.py:110-130."""
    mat = _materials()[mat_i]
    Ek = _as_energy(Ek)
    const = np.pi * mat.DA[val] ** 2 * Parameters.kT / (Parameters.hbar_eVs * mat.vs ** 2 * mat.rho)
    return const * density_of_states(Ek, mat.mass[val]) * 1.5E9


def _polar_prefactor(Ek, mat, val, quantum, eps):
    k = _wavevector(Ek, mat, val)
    return Parameters.q ** 2 * quantum * k / (8 * np.pi * eps * Parameters.hbar * (Ek + 1E-6))


def POP_abs(Ek, mat_i, val, T):
    """Polar optical phonon absorption, scatter_.py:133-156."""
    mat = _materials()[mat_i]
    Ek = _as_energy(Ek)
    eps_p = 1 / (1 / mat.eps_inf - 1 / mat.eps_0)
    p_abs = Bose_Einstein(mat.E_phonon) * np.arcsinh(np.sqrt(Ek / mat.E_phonon))
    return _polar_prefactor(Ek, mat, val, mat.E_phonon, eps_p) * p_abs * 2


def POP_ems(Ek, mat_i, val, T):
    """Polar optical phonon emission, scatter_.py:159-182."""
    mat = _materials()[mat_i]
    Ek = _as_energy(Ek)
    eps_p = 1 / (1 / mat.eps_inf - 1 / mat.eps_0)
    p_ems = (Ek > mat.E_phonon) * (Bose_Einstein(mat.E_phonon) + 1) * \
        np.arcsinh(np.sqrt(abs(Ek / mat.E_phonon - 1)))
    return _polar_prefactor(Ek, mat, val, mat.E_phonon, eps_p) * p_ems * 2.25


def IMP(Ek, mat_i, val, T):
    """Ionized impurity (Brooks-Herring-like), scatter_.py:583-624."""
    mat = _materials()[mat_i]
    Ek = _as_energy(Ek)
    Ld = 1 / np.sqrt(Parameters.q * mat.dope * 1E2 ** 3 / (mat.eps_0 * Parameters.kT))
    k = _wavevector(Ek, mat, val)
    # The reference evaluates this per energy through np.vectorize, so its `k**2` is the SCALAR
    # power = libm pow(k, 2.0), which differs from the array fast path k*k in the last bit for
    # ~0.1 % of the arguments; Python's float ** is the same libm call.
    k2 = np.array([float(x) ** 2 for x in np.ravel(k)], dtype=np.float64).reshape(np.shape(k))
    return (2 * np.pi * mat.dope * Parameters.q ** 2 / (Parameters.hbar_eVs * mat.eps_0 ** 2)) * \
        density_of_states(Ek, mat.mass[val]) * (1 / Ld ** 2) * (1 / (1 / Ld ** 2 + 4 * k2)) * 1E2


def cc(Ek, mat_i, val, T):
    """Carrier-carrier (plasmon-like), scatter_.py:626-652."""
    mat = _materials()[mat_i]
    Ek = _as_energy(Ek)
    E_plasma = plasmon_freq(mat.dope, mat_i, val) * Parameters.hbar_eVs
    coeff = _polar_prefactor(Ek, mat, val, E_plasma, mat.eps_0)
    p_abs = Bose_Einstein(E_plasma) * np.arcsinh(np.sqrt(Ek / E_plasma))
    p_ems = (Ek > E_plasma) * (Bose_Einstein(E_plasma) + 1) * \
        np.arcsinh(np.sqrt(abs(Ek / E_plasma - 1)))
    return coeff * (p_abs + p_ems)


# ---------------------------------------------------------------------------------
# intervalley mechanisms: one generator, sixteen descriptors (scatter_.py:185-580)
#   name: (pair key, gap sign (+1: '+E_val', -1: '-E_val', 0: none), gated?, fudge)
# absorption: argument Ek + EP (+/-) E_val, occupation N;  emission: Ek - EP (+/-) E_val, N+1.
# The density of states always uses the INITIAL valley's mass (scatter_.py:206,307,454).
# ---------------------------------------------------------------------------------
_IV_TABLE = {
    "IV_GL_abs": ("GL", -1, True, 5E15),   "IV_GL_ems": ("GL", -1, True, 5E15),
    "IV_GX_abs": ("GX", -1, True, 1.5E14), "IV_GX_ems": ("GX", -1, True, 1.75E14),
    "IV_LG_abs": ("LG", +1, False, 4E12),  "IV_LG_ems": ("LG", +1, False, 9E12),
    "IV_LL_abs": ("LL", 0, False, 9E13),   "IV_LL_ems": ("LL", 0, True, 9E13),
    "IV_LX_abs": ("LX", -1, True, 5E13),   "IV_LX_ems": ("LX", -1, True, 4E13),
    "IV_XG_abs": ("XG", +1, False, 5E10),  "IV_XG_ems": ("XG", +1, False, 4E10),
    "IV_XL_abs": ("XL", +1, False, 2E13),  "IV_XL_ems": ("XL", +1, False, 1.6E13),
    "IV_XX_abs": ("XX", 0, False, 5E12),   "IV_XX_ems": ("XX", 0, False, 4E12),
}


def _make_intervalley(name):
    key, gap_sign, gated, fudge = _IV_TABLE[name]
    emission = name.endswith("_ems")
    v_from, v_to = _VALLEY_LETTER.index(key[0]), _VALLEY_LETTER.index(key[1])

    def rate(Ek, mat_i, val, T):
        mat = _materials()[mat_i]
        Ek = _as_energy(Ek)
        EP = mat.EP[key]
        const = (np.pi * (mat.defpot[key]) ** 2 * mat.B[v_to]) / (mat.rho * EP / Parameters.hbar_eVs)
        arg = Ek - EP if emission else Ek + EP
        if gap_sign > 0:
            arg = arg + mat.E_val[key]
        elif gap_sign < 0:
            arg = arg - mat.E_val[key]
        occupation = Bose_Einstein(EP) + 1 if emission else Bose_Einstein(EP)
        p = density_of_states(abs(arg), mat.mass[v_from]) * occupation
        if gated:
            if gap_sign == 0:
                threshold = EP                                  # IV_LL_ems, scatter_.py:379
            elif emission:
                threshold = mat.E_val[key] + EP                 # e.g. scatter_.py:231
            else:
                threshold = mat.E_val[key] - EP                 # e.g. scatter_.py:205
            p = (Ek > threshold) * density_of_states(abs(arg), mat.mass[v_from]) * occupation
        return const * (p) * fudge

    rate.__name__ = name
    rate.__doc__ = "Intervalley %s %s, initial valley %d -> %d (generated; scatter_.py:185-580)." % (
        key, "emission" if emission else "absorption", v_from, v_to)
    return rate


for _n in _IV_TABLE:
    globals()[_n] = _make_intervalley(_n)

_REAL_MECHS = {
    0: ("ADP", "POP_abs", "POP_ems", "IV_GL_abs", "IV_GL_ems", "IV_GX_abs", "IV_GX_ems", "IMP", "cc"),
    1: ("ADP", "POP_abs", "POP_ems", "IV_LG_abs", "IV_LG_ems", "IV_LL_abs", "IV_LL_ems",
        "IV_LX_abs", "IV_LX_ems", "IMP", "cc"),
    2: ("ADP", "POP_abs", "POP_ems", "IV_XG_abs", "IV_XG_ems", "IV_XL_abs", "IV_XL_ems",
        "IV_XX_abs", "IV_XX_ems", "IMP", "cc"),
}


def _total(valley):
    def total(Ek, mat_i, val, T):
        names = _REAL_MECHS[valley]
        acc = globals()[names[0]](Ek, mat_i, val, T)
        for n in names[1:]:
            acc = acc + globals()[n](Ek, mat_i, val, T)
        return acc
    return total


def _self(valley):
    tot = _total(valley)

    def self_scatter(Ek, mat_i, val, T):
        return T - tot(Ek, mat_i, val, T)
    return self_scatter


total_scatter_G, total_scatter_L, total_scatter_X = _total(0), _total(1), _total(2)   # :654-726
self_scatter_G, self_scatter_L, self_scatter_X = _self(0), _self(1), _self(2)         # :728-792


# ---------------------------------------------------------------------------------
# polar-angle samplers (scatter_.py:795-889).  On the hot path these run inside the CUDA
# kernel; the host versions exist for API compatibility (they draw from np.random like
# the reference and return the ANGLE in radians despite their names).
# ---------------------------------------------------------------------------------
def iso_costheta(Ek, mat_i):
    return np.arccos(1 - 2 * np.random.uniform(low=0, high=1))


def pop_abs_costheta(Ek, mat_i):
    Ep = _materials()[mat_i].E_phonon
    eta = 2 * np.sqrt(Ek * (Ek + Ep)) / (np.sqrt(Ek) - np.sqrt(Ek + Ep)) ** 2
    r = np.random.uniform(0, 1)
    return np.arccos(((1 + eta) - (1 + 2 * eta) ** r) / eta)


def pop_ems_costheta(Ek, mat_i):
    Ep = _materials()[mat_i].E_phonon
    eta = 2 * np.sqrt(Ek * (Ek - Ep)) / (np.sqrt(Ek) - np.sqrt(Ek - Ep)) ** 2
    r = np.random.uniform(0, 1)
    return np.arccos(((1 + eta) - (1 + 2 * eta) ** r) / eta)


def ion_eb(mat_i):
    """The energy scale of ``ion_costheta`` (scatter_.py:868-869), a constant per material."""
    mat = _materials()[mat_i]
    b = (1 / (2 * mat.dope)) ** (1 / 3)
    return Parameters.q / (2 * mat.eps_0 * b)


def ion_costheta(Ek, mat_i):
    alpha = 2 * np.arctan(ion_eb(mat_i) / (Ek + 1E-6))
    r = np.random.uniform(low=0, high=1)
    return np.arccos(1 - (1 - np.cos(alpha)) / (1 - r * (1 - 0.5 * (1 - np.cos(alpha)))))


def self_scat_angle(Ek, mat_i):
    return 0


SAMPLER_OF = {self_scat_angle: SAMPLER_SELF, iso_costheta: SAMPLER_ISO,
              pop_abs_costheta: SAMPLER_POP_ABS, pop_ems_costheta: SAMPLER_POP_EMS,
              ion_costheta: SAMPLER_ION}


def scatter_mech(mat_i):
    """Per-valley ordered mechanism dictionaries (scatter_.py:892-951).

    Column order == selection order of the cumulative table.  ``dE`` is the energy added
    to the carrier, ``trans`` the final valley, ``polar`` the polar-angle sampler."""
    mat = _materials()[mat_i]
    Eph = mat.E_phonon
    n_val = len(mat.mass.keys())
    E_plasma = [plasmon_freq(mat.dope, mat_i, v) * Parameters.hbar_eVs for v in range(n_val)]
    g = globals()

    def entry(fn, dE, polar, trans):
        return {"mech": fn, "dE": dE, "polar": polar, "trans": trans}

    def intervalley(src):
        out = {}
        for dst in _VALLEY_LETTER:
            key = src + dst
            if key not in mat.EP:
                continue
            v_to = _VALLEY_LETTER.index(dst)
            if key in mat.E_val:
                # leaving Gamma or dropping L->X costs the gap; climbing back returns it
                gap = -mat.E_val[key] if _IV_TABLE["IV_%s_abs" % key][1] < 0 else mat.E_val[key]
                dE_abs, dE_ems = mat.EP[key] + gap, -mat.EP[key] + gap
            else:
                dE_abs, dE_ems = mat.EP[key], -mat.EP[key]
            out["IV_%s Absorption" % key] = entry(g["IV_%s_abs" % key], dE_abs, iso_costheta, v_to)
            out["IV_%s Emission" % key] = entry(g["IV_%s_ems" % key], dE_ems, iso_costheta, v_to)
        return out

    valleys = []
    for v, selfscat in enumerate((self_scatter_G, self_scatter_L, self_scatter_X)):
        d = {"ADP": entry(ADP, Eph, iso_costheta, v),
             "POP Absorption": entry(POP_abs, Eph, pop_abs_costheta, v),
             "POP Emission": entry(POP_ems, -Eph, pop_ems_costheta, v)}
        d.update(intervalley(_VALLEY_LETTER[v]))
        d["Ionized Impurity"] = entry(IMP, 0, ion_costheta, v)
        d["Carrier-Carrier"] = entry(cc, E_plasma[v], ion_costheta, v)
        d["Self-Scattering"] = entry(selfscat, 0, self_scat_angle, v)
        valleys.append(d)
    arr_ = np.array(valleys)
    return arr_[:n_val]


def calc_scatter_table(dfEk, mat_i):
    """Cumulative probability tables, one DataFrame ``(n_energy, n_mech)`` per valley
    (scatter_.py:954-984): row-wise cumsum of the rates divided by the self-scattering
    ceiling ``tot_scat``; the last column is 1."""
    import pandas as pd

    ek = dfEk["Ek"].to_numpy() if hasattr(dfEk, "columns") else np.asarray(dfEk, dtype=np.float64)
    ST = _materials()[mat_i].tot_scat
    out = []
    for v, scat in enumerate(scatter_mech(mat_i)):
        with np.errstate(all="ignore"):
            cols = {key: np.asarray(m["mech"](ek, mat_i, v, ST), dtype=np.float64)
                    for key, m in scat.items()}
        out.append(pd.DataFrame(cols).cumsum(1) / ST)
    return out


def mechanism_arrays(mat_i):
    """Flat per-valley metadata for the device: ``[(names, dE[], trans[], sampler[]), ...]``."""
    out = []
    for scat in scatter_mech(mat_i):
        names = list(scat.keys())
        out.append((names,
                    np.array([float(scat[k]["dE"]) for k in names], dtype=np.float64),
                    np.array([int(scat[k]["trans"]) for k in names], dtype=np.int32),
                    np.array([SAMPLER_OF[scat[k]["polar"]] for k in names], dtype=np.int32)))
    return out


# ---------------------------------------------------------------------------------
# per-event API with the reference's signatures, executed on the GPU
# ---------------------------------------------------------------------------------
def calc_energy(k, val, mat, dfEk):
    """(E_nonparabolic, nearest grid index), scatter_.py:19-54.  Runs ``emc_calc_energy``."""
    from . import main_
    return main_._calc_energy_gpu(k, val, mat, dfEk, check_zero=True)


def scatter(coords, scat_table, val, dfE, mat_i):
    """One scattering event for one particle (scatter_.py:986-1099), on the GPU.

    Keeps the reference's global-stream behaviour: draws r2, then the azimuth, then (for
    every mechanism except self-scattering) the polar uniform from ``np.random.uniform``
    in that order, so a seeded script consumes the stream exactly like the reference.
    ``coords`` is updated in place and returned together with the new valley."""
    from . import main_
    return main_._scatter_gpu(coords, scat_table, val, dfE, mat_i)
