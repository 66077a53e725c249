"""Source datasets: the host-side mirror of at3d/source.py (``solar`` :14, ``thermal`` :81, ``combined`` :143).

Same arguments, checks and variables (``wavelength, solarflux, solarmu, solaraz, srctype, units, wavenumber, skyrad``) as
plain mappings.  Volume point sources (``VolumeSource`` :212) are not on the path: passing one raises."""
import numpy as np
from ._dataset import Dataset


def _solar_mu(solarmu):
    solarmu = -1 * np.abs(solarmu)
    if not (-1.0 <= solarmu < 0.0):
        raise ValueError("solarmu must be in the range -1.0 <= solarmu < 0.0 not '{}'. "
                         "The SHDOM convention for solar direction is that it points"
                         "in the direction of the propagation of radiance.".format(solarmu))
    return solarmu


def _dataset(volume_source, **variables):
    if volume_source is not None:
        raise NotImplementedError('volume sources (at3d.source.VolumeSource) are not implemented')
    return Dataset(name='solar_source', wavenumber=np.array([10000, 10001]), **variables)


def solar(wavelength, solarmu, solar_azimuth, solarflux=1.0, skyrad=0.0, volume_source=None):
    """Collimated solar beam; `solar_azimuth` in degrees, `skyrad` an isotropic radiance from above."""
    return _dataset(volume_source, wavelength=wavelength, solarflux=solarflux, solarmu=_solar_mu(solarmu),
                    solaraz=np.deg2rad(solar_azimuth), srctype='S', units='R', skyrad=np.atleast_3d(skyrad))


def thermal(wavelength, skyrad=0.0, units='radiance', volume_source=None):
    """Thermal emission; `skyrad` is the brightness temperature of the radiance from above."""
    if units not in ('radiance', 'brightness_temperature'):
        raise ValueError("`units` should be either 'radiance' or 'brightness_temperature'.")
    return _dataset(volume_source, wavelength=wavelength, solarflux=0.0, solarmu=-0.5, solaraz=0.0, srctype='T',
                    units='R' if units == 'radiance' else 'T', skyrad=skyrad)


def combined(wavelength, solarmu, solar_azimuth, solarflux=1.0, skyrad=0.0, volume_source=None):
    """Solar beam and thermal emission together."""
    return _dataset(volume_source, wavelength=wavelength, solarflux=solarflux, solarmu=_solar_mu(solarmu),
                    solaraz=np.deg2rad(solar_azimuth), srctype='B', units='R', skyrad=skyrad)
