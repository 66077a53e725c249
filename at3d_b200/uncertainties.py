"""Measurement uncertainty models: the host-side mirror of at3d/uncertainties.py (``Uncertainty`` :20, ``NullUncertainty``
:100, ``RadiometricUncertainty`` :126, ``TabulatedRadiometricUncertainty`` :284, ``ResearchScanningPolarimeter`` :337,
``TandemStereoCamera`` :348).

They fill the ``uncertainties`` variable of a sensor -- the inverse error covariance ``[num_uncertainty, num_uncertainty,
npixels]`` that COMPUTE_ADJOINT_WEIGHTS / UPDATE_COSTFUNCTION of the gradient read -- and draw measurement noise.
Same constructor arguments, formulas and order of the NumPy global random draws as the reference, on plain-mapping sensors
(``sensor['I']`` per pixel, ``sensor['stokes']``).
"""
import numpy as np

_STOKES = ('I', 'Q', 'U', 'V')


def _get(sensor, name):
    x = sensor[name]
    return np.asarray(getattr(x, 'data', x))


class Uncertainty:
    """Base class: which cost function the inverse covariance is for ('L2': 4 x 4 Stokes, 'LL': 2 x 2 log-radiance / DoLP)."""

    def __init__(self, cost_function):
        self._valid_cost_functions = ('L2', 'LL')
        if cost_function not in self._valid_cost_functions:
            raise NotImplementedError("`cost_function` '{}' is not supported for this Uncertainty type. "
                                      "Valid values are '{}'".format(cost_function, self._valid_cost_functions))
        self._cost_function = cost_function
        self._num_uncertainty = 4 if cost_function == 'L2' else 2

    def calculate_uncertainties(self, sensor):
        sensor['uncertainties'] = self._process_uncertainties(sensor)

    def add_noise(self, sensor, noise=True):
        perturbed_stokes = self._process_noise(sensor)
        for i, has_stokes in enumerate(_get(sensor, 'stokes')):
            if has_stokes:
                if _STOKES[i] not in sensor:
                    raise KeyError("Stokes component '{}' is not found in sensor even though it is an observable. Noise "
                                   "perturbations for this observable cannot be generated.".format(_STOKES[i]))
                sensor[_STOKES[i]] = np.array(perturbed_stokes[i], dtype=_get(sensor, _STOKES[i]).dtype)

    @property
    def cost_function(self):
        return self._cost_function

    @property
    def num_uncertainty(self):
        return self._num_uncertainty

    @property
    def valid_cost_functions(self):
        return self._valid_cost_functions


class NullUncertainty(Uncertainty):
    """Uniform weights `scaling_factor` in every entry (the reference fills the whole matrix, :114-116); no noise."""

    def __init__(self, cost_function, scaling_factor=1.0):
        super().__init__(cost_function)
        self._scaling_factor = scaling_factor

    def _process_uncertainties(self, sensor):
        npixels = _get(sensor, 'npixels').size if 'npixels' in sensor and np.ndim(_get(sensor, 'npixels')) > 0 \
            else _get(sensor, 'cam_mu').size if 'cam_mu' in sensor else int(_get(sensor, 'pixel_index').max()) + 1
        return self._scaling_factor * np.ones((self._num_uncertainty, self._num_uncertainty, npixels))

    def add_noise(self, sensor):
        raise ValueError("{} cannot be used to generate measurement noise. Please assign another uncertainty "
                         "model.".format(type(self)))


class RadiometricUncertainty(Uncertainty):
    """Intensity-only radiometric noise: signal-to-noise curve `snr_func(radiance)` with a floor `sigma_floor`, a
    camera-to-camera and an absolute calibration uncertainty (fractions) (:126-282)."""

    def __init__(self, cost_function, snr_func, sigma_floor, absolute_calibration_uncertainty=0.0,
                 camera_to_camera_calibration_uncertainty=0.0, seed=None):
        super().__init__(cost_function)
        self._snr_func, self._floor = snr_func, sigma_floor
        self._camera_sigma, self._absolute_sigma = camera_to_camera_calibration_uncertainty, absolute_calibration_uncertainty
        if seed is not None:
            np.random.seed(seed)
        # one draw per instrument, shared by all of its images (and drawn even when the uncertainty is zero, so that
        # the global random stream advances as it does in the reference)
        self._absolute_bias = np.random.normal(loc=0.0, scale=absolute_calibration_uncertainty)

    def noise_curve(self, reflectance):
        return self._snr_func(np.atleast_1d(reflectance))

    def _noise_variance(self, radiance):
        """(radiance / SNR)^2, not below the floor; the floor where the curve gives no signal."""
        snr = self.noise_curve(radiance)
        sigma = np.maximum(radiance / snr, self._floor)
        sigma[np.where(snr == 0.0)] = self._floor
        return sigma ** 2

    def _process_noise(self, sensor, seed=None, camera_cal=True, noise=True, absolute_cal=True):
        """Perturbed Stokes vector [4, npixels]: only I is perturbed (:208-243)."""
        radiance = _get(sensor, 'I')
        if seed is not None:
            np.random.seed(seed)
        if noise:
            # Poisson counts with the mean and the variance of the signal, scaled back to radiance
            variance = self._noise_variance(radiance)
            counts = np.random.poisson((radiance ** 2) / variance)
            radiance = np.sqrt(counts * variance)
        else:
            radiance = np.array(radiance)
        if camera_cal:
            radiance *= np.random.normal(loc=1.0, scale=self._camera_sigma)
        if absolute_cal:
            radiance *= (1.0 + self._absolute_bias)
        perturbed = np.zeros((4, radiance.size))
        perturbed[0] = radiance
        return perturbed

    def _process_uncertainties(self, sensor, camera_cal=True, noise=True, absolute_cal=True):
        """Inverse variance of I in entry [0, 0], ones elsewhere (:245-282)."""
        radiance = _get(sensor, 'I')
        terms = ([self._noise_variance(radiance)] if noise else []) \
            + ([(self._camera_sigma * radiance) ** 2] if camera_cal else []) \
            + ([(self._absolute_bias * radiance) ** 2] if absolute_cal else [])
        inverse_variance = 1.0 / sum(terms) if terms else np.ones(radiance.shape)
        out = np.ones((self._num_uncertainty, self._num_uncertainty, inverse_variance.size))
        out[0, 0] = inverse_variance
        return out


class TabulatedRadiometricUncertainty(RadiometricUncertainty):
    """SNR tabulated against reflectance: cubic spline inside the table, ``a + b sqrt(x)`` fitted above it, constant noise
    below it (:284-331).  As in the reference the calibration arguments are accepted and NOT passed on (:331)."""

    def __init__(self, cost_function, reflectance_values, SNR_values, absolute_calibration_uncertainty=0.0,
                 camera_to_camera_calibration_uncertainty=0.0, seed=None):
        from scipy.interpolate import CubicSpline
        from scipy.optimize import curve_fit
        self._table = np.asarray(reflectance_values)
        self._spline = CubicSpline(reflectance_values, SNR_values, extrapolate=True)
        self._fit, _ = curve_fit(self._sqrt_law, reflectance_values, SNR_values, p0=[1e-5, 10])
        self._noise_at_table_start = (reflectance_values / SNR_values)[0]
        super().__init__(cost_function, self._tabulated_snr, self._noise_at_table_start)

    @staticmethod
    def _sqrt_law(x, a, b):
        return a + b * np.sqrt(x)

    def _tabulated_snr(self, x):
        reflectance = np.atleast_1d(x)
        snr = self._spline(reflectance)
        above = np.where(reflectance > self._table.max())
        snr[above] = self._sqrt_law(reflectance[above], *self._fit)
        below = np.where(reflectance < self._table.min())
        snr[below] = reflectance[below] / self._noise_at_table_start
        return snr


class ResearchScanningPolarimeter(RadiometricUncertainty):
    """The reference passes the number 2e-5 where a callable SNR curve is expected (:343-345): kept, so noise_curve
    raises exactly as it does there."""

    def __init__(self, cost_function):
        super().__init__(cost_function, 2e-5, 1e-7, camera_to_camera_calibration_uncertainty=0.0,
                         absolute_calibration_uncertainty=np.sqrt(0.015))


class TandemStereoCamera(TabulatedRadiometricUncertainty):
    def __init__(self, cost_function, camera_to_camera_calibration_uncertainty=0.01, absolute_calibration_uncertainty=0.03):
        super().__init__(cost_function, np.array([0.01, 0.05, 0.1, 0.5, 1.0, 1.3]) / np.pi,
                         np.array([87.0, 201.0, 285.0, 639.0, 904.0, 1031.0]),
                         absolute_calibration_uncertainty=absolute_calibration_uncertainty,
                         camera_to_camera_calibration_uncertainty=camera_to_camera_calibration_uncertainty)
