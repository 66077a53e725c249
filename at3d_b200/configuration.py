"""Numerical parameters: the host-side mirror of at3d/configuration.py (``make_config_data`` :55, ``get_config`` :172,
``make_config`` :11): the same names and default values, as a plain mapping ``name -> value``; a JSON file in the
reference's layout (``name -> {'default_value': ..., 'description': ...}``) is read and written the same way."""
import json
from ._dataset import Dataset

_DEFAULTS = (
    ('x_boundary_condition', 'open'), ('y_boundary_condition', 'open'), ('num_mu_bins', 16), ('num_phi_bins', 32),
    ('split_accuracy', 0.03), ('deltam', True), ('spherical_harmonics_accuracy', 0.0), ('acceleration_flag', True),
    ('solution_accuracy', 0.0001), ('max_total_mb', 3000.0), ('adapt_grid_factor', 5), ('num_sh_term_factor', 1),
    ('cell_to_point_ratio', 1.5), ('high_order_radiance', False), ('ip_flag', 0), ('iterfixsh', 30), ('tautol', 0.1),
    ('angle_set', 2), ('transcut', 1e-5), ('transmin', 1.0))


def make_config_data(**parameters):
    """``name -> {'default_value': value}`` with the reference's defaults, overridden by keyword."""
    unknown = set(parameters) - {k for k, _ in _DEFAULTS}
    if unknown:
        raise TypeError('unknown numerical parameters: {}'.format(sorted(unknown)))
    return {k: {'default_value': parameters.get(k, v)} for k, v in _DEFAULTS}


def make_config(config_file_name, **parameters):
    with open(config_file_name, 'w') as f:
        json.dump(make_config_data(**parameters), f, indent=4)


def get_config(config_file_name=None):
    if config_file_name is None:
        configuration = make_config_data()
    else:
        with open(config_file_name, 'r') as f:
            configuration = json.load(f)
    return Dataset((k, a['default_value']) for k, a in configuration.items())
