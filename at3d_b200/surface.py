"""Surface datasets: the host-side mirror of at3d/surface.py (``lambertian`` :28, ``wave_fresnel`` :106, ``diner`` :204,
``ocean_unpolarized`` :308, ``RPV_unpolarized`` :404, ``_make_surface_dataset`` :565) with ``PREP_SURFACE``
(src/surface.f:2-288) restated on the host: the same arguments and checks, the same variables (``sfctype, gndalbedo,
gndtemp, maxsfcpars, nxsfc, nysfc, delxsfc, delysfc, nsfcpar, sfcparms``), ``sfcparms`` laid out as
``[nsfcpar, nxsfc+1, nysfc+1]`` (parameter fastest) with the periodic edge copies, ``gndtemp`` / ``gndalbedo`` the float32
running means PREP_SURFACE returns.  Not provided: the Ross-Li surface ('VM', ``ross_li_thick_sparse`` :492; ``RTE``
refuses it) and surface point sources (``SurfaceSource`` :671)."""
import numpy as np
from ._dataset import Dataset

# SFCTYPE, number of parameters (temperature first), whether the second parameter is averaged into GNDALBEDO
_TYPES = {'variable_lambertian': ('VL', 2, True), 'wave_fresnel': ('VW', 4, False), 'diner': ('VD', 6, True),
          'ocean_unpolarized': ('VO', 3, False), 'rpv_unpolarized': ('VR', 4, True)}


def _no_source(surface_source):
    if surface_source is not None:
        raise NotImplementedError('surface sources (at3d.surface.SurfaceSource) are not implemented')


def prep_surface(sfctype, parms_in):
    """PREP_SURFACE (src/surface.f:2-288) for `parms_in` [npar, nxsfc, nysfc] (temperature first): returns
    ``(sfcparms[npar * (nxsfc+1) * (nysfc+1)], gndtemp, gndalbedo)``."""
    parms_in = np.asarray(parms_in, np.float32)
    npar, nxs, nys = parms_in.shape
    if sfctype == 'VL' and (np.any(parms_in[1] < 0.0) or np.any(parms_in[1] > 1.0)):
        raise ValueError('PREP_SURFACE: Illegal surface albedo')
    parms = np.zeros((npar, nxs + 1, nys + 1), np.float32, order='F')
    parms[:, :nxs, :nys] = parms_in
    parms[:, :nxs, nys] = parms[:, :nxs, 0]                # the row beyond the last repeats the first (periodic surface)
    parms[:, nxs, :] = parms[:, 0, :]
    # means accumulated in float32 in the order of the points (x fastest), as the Fortran loop does
    def mean(field):
        return np.float32(np.cumsum(field.ravel(order='F'), dtype=np.float32)[-1] / np.float32(field.size))
    has_albedo = {t: a for t, _, a in _TYPES.values()}[sfctype]
    return parms.ravel(order='F'), mean(parms_in[0]), (mean(parms_in[1]) if has_albedo else np.float32(0.0))


def _make_surface_dataset(surface_type, ground_temperature, delx, dely, surface_source, **kwargs):
    _no_source(surface_source)
    sfctype, npar, _ = _TYPES[surface_type]
    nxsfc, nysfc = ground_temperature.shape
    parms_in = np.stack([ground_temperature] + list(kwargs.values()), axis=0)
    assert parms_in.shape[0] == npar
    sfcparms, gndtemp, gndalbedo = prep_surface(sfctype, parms_in)
    return Dataset(name=surface_type, sfctype=sfctype, gndalbedo=gndalbedo, gndtemp=gndtemp, maxsfcpars=npar,
                   nxsfc=nxsfc, nysfc=nysfc, delxsfc=delx, delysfc=dely, nsfcpar=npar, sfcparms=sfcparms)


def _variable(surface_type, ground_temperature, delx, dely, surface_source, **parameters):
    """The argument handling the reference repeats in every variable-surface factory (:143-176, :249-281, ...)."""
    parameters = {k: np.atleast_2d(v) for k, v in parameters.items()}
    first = next(iter(parameters.values()))
    if any(p.shape != first.shape for p in parameters.values()):
        raise ValueError('All surface brdf parameters must have the same shape.')
    ground_temperature = np.atleast_2d(ground_temperature)
    if ground_temperature.size != 1:
        raise ValueError('ground temperature must have a compatible shape')
    # np.full_like, as the reference: the temperature field takes the dtype of the reference parameter
    ground_temperature = np.full_like(parameters[_LIKE[surface_type]], fill_value=ground_temperature[0, 0])
    if first.size == 1:
        delx = 0.02 if delx is None else delx
        dely = 0.02 if dely is None else dely
    elif dely is None or delx is None:
        raise ValueError('dely and delx must be defined for variable surface parameters.')
    return _make_surface_dataset(surface_type, ground_temperature, delx, dely, surface_source, **parameters)


_LIKE = {'wave_fresnel': 'real_refractive_index', 'diner': 'A', 'ocean_unpolarized': 'pigmentation', 'rpv_unpolarized': 'K'}


def lambertian(albedo, ground_temperature=298.15, delx=None, dely=None, surface_source=None):
    """Lambertian surface: fixed ('FL') for a scalar albedo, variable ('VL') for a 2-D albedo map with spacing
    `delx`, `dely`."""
    if np.any(np.asarray(albedo) > 1.0) or np.any(np.asarray(albedo) < 0.0):
        raise ValueError("surface albedo should be in [0, 1] not '{}'".format(albedo))
    ground_temperature = np.atleast_2d(ground_temperature)
    albedo = np.atleast_2d(albedo)
    if albedo.size == 1 and ground_temperature.size == 1:
        _no_source(surface_source)
        return Dataset(name='fixed_lambertian_surface', sfctype='FL', gndalbedo=albedo[0, 0],
                       gndtemp=ground_temperature[0, 0], maxsfcpars=4, nxsfc=0, nysfc=0, delxsfc=0, delysfc=0, nsfcpar=1,
                       sfcparms=np.zeros(0, np.float32))
    if ground_temperature.size != 1:
        raise ValueError('ground temperature must have a compatible shape.')
    if dely is None or delx is None:
        raise ValueError('dely and delx must be defined for variable surface parameters.')
    ground_temperature = np.full_like(albedo, fill_value=ground_temperature[0, 0])
    return _make_surface_dataset('variable_lambertian', ground_temperature, delx, dely, surface_source, albedo=albedo)


def wave_fresnel(real_refractive_index, imaginary_refractive_index, surface_wind_speed, ground_temperature=298.15,
                 delx=None, dely=None, surface_source=None):
    """Polarized Fresnel reflection from a wind-roughened water surface ('VW'; wind speed in m/s)."""
    return _variable('wave_fresnel', ground_temperature, delx, dely, surface_source,
                     real_refractive_index=real_refractive_index, imaginary_refractive_index=imaginary_refractive_index,
                     surface_wind_speed=surface_wind_speed)


def diner(A, K, B, ZETA, SIGMA, ground_temperature=298.15, delx=None, dely=None, surface_source=None):
    """Diner et al. polarized modified-RPV surface ('VD')."""
    return _variable('diner', ground_temperature, delx, dely, surface_source, A=A, K=K, B=B, ZETA=ZETA, SIGMA=SIGMA)


def ocean_unpolarized(surface_wind_speed, pigmentation, ground_temperature=298.15, delx=None, dely=None,
                      surface_source=None):
    """Unpolarized ocean BRDF ('VO'; wind speed in m/s, pigment concentration in mg/m^3)."""
    return _variable('ocean_unpolarized', ground_temperature, delx, dely, surface_source,
                     surface_wind_speed=surface_wind_speed, pigmentation=pigmentation)


def RPV_unpolarized(RHO0, K, THETA, ground_temperature=298.15, delx=None, dely=None, surface_source=None):
    """Rahman-Pinty-Verstraete land surface ('VR')."""
    return _variable('rpv_unpolarized', ground_temperature, delx, dely, surface_source, RHO0=RHO0, K=K, THETA=THETA)
