"""Sensor datasets: the host-side mirror of at3d/sensor.py -- ``make_sensor_dataset`` (:31), the sub-pixel ray generators
``stochastic`` / ``uniform`` / ``gaussian`` (:109-186), ``orthographic_projection`` (:188), ``perspective_projection`` (:367)
and ``domaintop_projection`` (:558).

Same arguments, checks, pixel order (x fastest within an image), variables (``wavelength, stokes[4], cam_x/y/z/mu/phi,
ray_x/y/z/mu/phi, ray_weight, pixel_index, use_subpixel_rays, image_shape, bounding_box``) and ``attrs`` as the
reference, with the same floating-point steps (float32 camera position and intrinsics in the perspective projection,
float64 angles), as plain mappings (at3d_b200/_dataset.py).  These are the ray lists RENDER and the gradient consume:
``RTE.integrate_to_sensor`` / ``SensorsDict.add_sensor`` take them as they are.
"""
import inspect
import itertools
import numpy as np
from ._dataset import Dataset

_RAY_FROM_CAM = (('ray_mu', 'cam_mu'), ('ray_phi', 'cam_phi'), ('ray_x', 'cam_x'), ('ray_y', 'cam_y'), ('ray_z', 'cam_z'))


def make_sensor_dataset(x, y, z, mu, phi, stokes, wavelength, fill_ray_variables=False):
    """Pixel positions and directions (direction of propagation: `mu` cosine of zenith, `phi` azimuth in radians) of one
    image (:31-107)."""
    x, y, z, mu, phi, wavelength = (np.asarray(a) for a in (x, y, z, mu, phi, wavelength))
    stokes = np.atleast_1d(stokes)
    for totest, name in zip((x, y, z, mu, phi), ('x', 'y', 'z', 'mu', 'phi')):
        if totest.ndim != 1:
            raise ValueError("'{}' should be a 1-D np.ndarray".format(name))
    if not all(x.size == a.size for a in (y, z, mu, phi)):
        raise ValueError("All of x, y, z, mu, phi should have the same size.")
    if not np.all(z >= 0.0):
        raise ValueError("All altitudes (z) must be >= 0.0")
    if np.any(mu == 0.0):
        raise ValueError("'mu' values of 0.0 are not allowed.")
    for component in stokes:
        if component not in ('I', 'Q', 'U', 'V'):
            raise ValueError("Valid Stokes components are 'I', 'Q', 'U', 'V' not '{}'".format(component))
    dataset = Dataset(wavelength=wavelength, stokes=np.array([c in stokes for c in ('I', 'Q', 'U', 'V')]),
                      cam_x=x.astype(np.float64), cam_y=y.astype(np.float64), cam_z=z.astype(np.float64),
                      cam_mu=mu.copy(), cam_phi=phi.copy())
    return _add_null_subpixel_rays(dataset) if fill_ray_variables else dataset


# ---- sub-pixel ray generators: (position perturbations in [-1, 1] of half a pixel, weights) per image dimension ----
def stochastic(npixels, nrays, seed=None):
    """`nrays` uniformly random positions per pixel, equal weights (:109-145)."""
    if seed is not None:
        np.random.seed(seed)
    return np.random.uniform(low=-1.0, high=1.0, size=(npixels, nrays)), np.ones((npixels, nrays)) / nrays


def uniform(npixels, nrays):
    """`nrays` equally spaced positions, the same for every pixel (:147-150)."""
    return np.linspace(-1.0 + 1 / nrays, 1.0 - 1 / nrays, nrays), np.ones((npixels, nrays)) / nrays


def gaussian(npixels, degree):
    """Gauss-Legendre nodes and normalised weights of `degree` (:152-186)."""
    nodes, weights = np.polynomial.legendre.leggauss(degree)
    return np.tile(nodes, (npixels, 1)), np.tile(weights, (npixels, 1)) / np.sum(weights)


def _parse_sub_pixel_ray_args(sub_pixel_ray_args):
    """(method, kwargs for x, kwargs for y): a tuple value gives the two dimensions separately (:624-659)."""
    method = sub_pixel_ray_args['method']
    try:
        parameters = inspect.signature(method).parameters
    except TypeError as err:
        raise TypeError("sub_pixel_ray_args 'method' must be a callable object not '{}' of type '{}'".format(
            method, type(method))) from err
    kwargs_x, kwargs_y = {}, {}
    for name, value in sub_pixel_ray_args.items():
        if name == 'method':
            continue
        if name not in parameters:
            raise KeyError("Invalid kwarg '{}' passed to the sub_pixel_ray_args['method'] callable. '{}'".format(
                name, method.__name__))
        kwargs_x[name], kwargs_y[name] = value if isinstance(value, tuple) else (value, value)
    return method, kwargs_x, kwargs_y


def _sub_pixel_offsets(npixels, sub_pixel_ray_args, pixel_x, pixel_y):
    """Offsets of every sub-pixel ray in the image plane, [npixels, nrays_x, nrays_y] each, and the ray weights."""
    method, kwargs_x, kwargs_y = _parse_sub_pixel_ray_args(sub_pixel_ray_args)
    pert_x, weights_x = method(npixels, **kwargs_x)
    pert_y, weights_y = method(npixels, **kwargs_y)
    shape = (npixels, weights_x.shape[-1], weights_y.shape[-1])
    off_x = np.broadcast_to(np.asarray(pert_x)[..., np.newaxis] * pixel_x / 2.0, shape)
    off_y = np.broadcast_to(np.asarray(pert_y)[..., np.newaxis, :] * pixel_y / 2.0, shape)
    weight = (weights_x[..., np.newaxis] * weights_y[..., np.newaxis, :]).ravel()
    return off_x, off_y, weight, shape


def _record_sub_pixel_args(sensor, sub_pixel_ray_args):
    # the reference also overwrites the caller's dictionary entry with the name (:335); here the argument is left alone
    for attribute, value in sub_pixel_ray_args.items():
        sensor.attrs['sub_pixel_ray_args_{}'.format(attribute)] = value.__name__ if attribute == 'method' else value


def _add_null_subpixel_rays(sensor):
    """One ray per pixel: the ray variables repeat the pixel variables (:661-683)."""
    for name in ('cam_mu', 'cam_phi', 'cam_x', 'cam_y', 'cam_z'):
        if name not in sensor:
            raise ValueError("'{}' is missing from sensor. This is not a valid sensor.".format(name))
    for ray, cam in _RAY_FROM_CAM:
        sensor[ray] = np.array(sensor[cam])
    npixels = len(sensor['cam_mu'])
    sensor['pixel_index'] = np.arange(npixels)
    sensor['ray_weight'] = np.ones(npixels)
    sensor['use_subpixel_rays'] = False
    return sensor


def _bounds(bounding_box):
    x, y, z = (np.asarray(getattr(bounding_box[k], 'data', bounding_box[k])) for k in ('x', 'y', 'z'))
    return x.min(), y.min(), z.min(), x.max(), y.max(), z.max()


def orthographic_projection(wavelength, bounding_box, x_resolution, y_resolution, azimuth, zenith, altitude='TOA',
                            stokes='I', sub_pixel_ray_args={'method': None}):
    """Parallel rays covering the projection of `bounding_box` (a grid: ``x, y, z``) onto the plane at `altitude`
    (:188-347).  `azimuth`, `zenith` in degrees (direction of propagation of the photons); resolutions in the units of the grid."""
    mu = np.cos(np.deg2rad(zenith))
    phi = np.deg2rad(azimuth)
    xmin, ymin, zmin, xmax, ymax, zmax = _bounds(bounding_box)
    altitude = zmax if isinstance(altitude, str) and altitude == 'TOA' else altitude
    # the corners of the box slide along the view direction onto the image plane
    alpha = np.sqrt(1 - mu ** 2) * np.cos(phi) / mu
    beta = np.sqrt(1 - mu ** 2) * np.sin(phi) / mu
    projection_matrix = np.array([[1, 0, -alpha, alpha * altitude], [0, 1, -beta, beta * altitude], [0, 0, 0, altitude]])
    corners = np.array(list(itertools.product([xmin, xmax], [ymin, ymax], [zmin, zmax]))).T
    projected = np.dot(projection_matrix, np.pad(corners, ((0, 1), (0, 0)), 'constant', constant_values=1))
    x_s, y_s = projected[:2, :].min(axis=1)
    x_e, y_e = projected[:2, :].max(axis=1)
    # padded so that there is a sample at (or past) the far edge (:265-271)
    x = np.arange(x_s, x_e + x_resolution, x_resolution)
    y = np.arange(y_s, y_e + y_resolution, y_resolution)
    image_shape = [x.size, y.size]
    x, y, z, mu, phi = (a.ravel() for a in np.meshgrid(x, y, altitude, mu, phi))
    sensor = make_sensor_dataset(x, y, z, mu, phi, stokes, wavelength)
    sensor['bounding_box'] = np.array([xmin, ymin, zmin, xmax, ymax, zmax])
    sensor['image_shape'] = np.array(image_shape)
    sensor.attrs = {'projection': 'Orthographic', 'altitude': altitude, 'x_resolution': x_resolution,
                    'y_resolution': y_resolution, 'projection_azimuth': azimuth, 'projection_zenith': zenith}
    if sub_pixel_ray_args['method'] is None:
        return _add_null_subpixel_rays(sensor)
    off_x, off_y, weight, shape = _sub_pixel_offsets(x.size, sub_pixel_ray_args, x_resolution, y_resolution)
    spread = lambda a: np.broadcast_to(a[:, np.newaxis, np.newaxis], shape).ravel()
    sensor['ray_mu'], sensor['ray_phi'] = spread(mu), spread(phi)
    sensor['ray_x'] = (x[:, np.newaxis, np.newaxis] + off_x).ravel()
    sensor['ray_y'] = (y[:, np.newaxis, np.newaxis] + off_y).ravel()
    sensor['ray_z'] = spread(z)
    sensor['pixel_index'] = np.repeat(np.arange(x.size), shape[1] * shape[2])
    sensor['ray_weight'] = weight
    sensor['use_subpixel_rays'] = True
    _record_sub_pixel_args(sensor, sub_pixel_ray_args)
    return sensor


def perspective_projection(wavelength, fov, x_resolution, y_resolution, position_vector, lookat_vector, up_vector,
                           stokes='I', sub_pixel_ray_args={'method': None}):
    """Pinhole camera at `position_vector` looking at `lookat_vector` (:367-556): `fov` degrees across the longer image
    side, `x_resolution` x `y_resolution` pixels."""
    def norm(v):
        return v / np.linalg.norm(v, axis=0)
    assert int(x_resolution) == x_resolution, "x_resolution is an integer >= 1"
    assert int(y_resolution) == y_resolution, "y_resolution is an integer >= 1"
    nx, ny = x_resolution, y_resolution
    position = np.array(position_vector, dtype=np.float32)
    lookat = np.array(lookat_vector, dtype=np.float32)
    zaxis = norm(lookat - position)
    xaxis = norm(np.cross(np.array(up_vector), zaxis))
    yaxis = np.cross(zaxis, xaxis)
    rotation_matrix = np.stack((xaxis, yaxis, zaxis), axis=1)
    extent = np.array([nx, ny]) / max(nx, ny)               # half-sizes of the normalised image plane
    dx, dy = 2 * extent[0] / nx, 2 * extent[1] / ny
    x_s, y_s, z_s = (a.ravel() for a in np.meshgrid(np.linspace(-extent[0] + dx / 2, extent[0] - dx / 2, nx),
                                                    np.linspace(-extent[1] + dy / 2, extent[1] - dy / 2, ny), 1.0))
    focal = 1.0 / np.tan(np.deg2rad(fov) / 2.0)
    k = np.array([[focal, 0, 0], [0, focal, 0], [0, 0, 1]], dtype=np.float32)
    inv_k = np.linalg.inv(k)

    def directions(xs, ys, zs):
        x_c, y_c, z_c = norm(np.matmul(rotation_matrix, np.matmul(inv_k, np.stack([xs, ys, zs]))))
        # propagation direction of the photons, towards the camera
        return -z_c.astype(np.float64), (np.arctan2(y_c, x_c) + np.pi).astype(np.float64)

    def at_camera(n):
        return tuple(np.full(n, position[i], dtype=np.float32) for i in range(3))

    mu, phi = directions(x_s, y_s, z_s)
    npix = nx * ny
    sensor = make_sensor_dataset(*at_camera(npix), mu, phi, stokes, wavelength)
    sensor['image_shape'] = np.array([nx, ny])
    sensor.attrs = {'projection': 'Perspective', 'fov_deg': fov, 'fov_x_deg': np.rad2deg(2 * np.arctan(extent[0] / focal)),
                    'fov_y_deg': np.rad2deg(2 * np.arctan(extent[1] / focal)), 'x_resolution': x_resolution,
                    'y_resolution': y_resolution, 'position': position, 'lookat': lookat,
                    'rotation_matrix': rotation_matrix.ravel(), 'sensor_to_camera_transform_matrix': k.ravel()}
    if sub_pixel_ray_args['method'] is None:
        return _add_null_subpixel_rays(sensor)
    off_x, off_y, weight, shape = _sub_pixel_offsets(npix, sub_pixel_ray_args, dx, dy)
    sensor['ray_mu'], sensor['ray_phi'] = directions((x_s[:, np.newaxis, np.newaxis] + off_x).ravel(),
                                                     (y_s[:, np.newaxis, np.newaxis] + off_y).ravel(),
                                                     np.broadcast_to(z_s[:, np.newaxis, np.newaxis], shape).ravel())
    sensor['ray_x'], sensor['ray_y'], sensor['ray_z'] = at_camera(weight.size)
    sensor['pixel_index'] = np.repeat(np.arange(npix), shape[1] * shape[2])
    sensor['ray_weight'] = weight
    sensor['use_subpixel_rays'] = True
    _record_sub_pixel_args(sensor, sub_pixel_ray_args)
    return sensor


def domaintop_projection(wavelength, bounding_box, x_resolution, y_resolution, azimuth, zenith, x_offset=0.0, y_offset=0.0,
                         stokes='I', sub_pixel_ray_args={'method': None}):
    """Parallel rays leaving the domain top on a regular grid of positions (a nadir orthographic grid, shifted by the
    offsets), all in the direction (`azimuth`, `zenith`) (:558-622)."""
    sensor = orthographic_projection(wavelength, bounding_box, x_resolution, y_resolution, 0.0, 0.0, altitude='TOA',
                                     stokes=stokes, sub_pixel_ray_args=sub_pixel_ray_args)
    for kind in ('cam', 'ray'):
        sensor[kind + '_x'] = sensor[kind + '_x'] + x_offset
        sensor[kind + '_y'] = sensor[kind + '_y'] + y_offset
        sensor[kind + '_mu'] = np.full_like(sensor[kind + '_mu'], np.cos(np.deg2rad(zenith)))
        sensor[kind + '_phi'] = np.full_like(sensor[kind + '_phi'], np.deg2rad(azimuth))
    sensor.attrs.update(projection='DomainTop', projection_azimuth=azimuth, projection_zenith=zenith)
    return sensor
