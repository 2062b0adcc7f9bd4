"""Import-time singleton of the ICP parameters, same attribute names as the reference's config/config.py:6-34
(max_radius, min_radius, max_height, min_height, voxel_size, radius_gd, max_nn_gd, radius_normals, max_nn,
distance_threshold), read from icp_parameters.yaml next to this module.  Extra, optional: the ICP convergence
criteria (Open3D defaults when the section is absent, as in the reference's yaml)."""
import os

import yaml

# attribute -> (yaml section, key, default); a default of ... marks a required entry
_FIELDS = {
    "max_radius": ("filter_by_radius", "max_radius", ...),
    "min_radius": ("filter_by_radius", "min_radius", ...),
    "max_height": ("filter_by_height", "max_height", ...),
    "min_height": ("filter_by_height", "min_height", ...),
    "voxel_size": ("down_sample", "voxel_size", None),
    "radius_gd": ("filter_ground_plane", "radius_normals", ...),
    "max_nn_gd": ("filter_ground_plane", "maximum_neighbors", ...),
    "radius_normals": ("normals", "radius_normals", ...),
    "max_nn": ("normals", "maximum_neighbors", ...),
    "distance_threshold": ("icp", "distance_threshold", ...),
    "relative_fitness": ("icp_criteria", "relative_fitness", 1e-6),
    "relative_rmse": ("icp_criteria", "relative_rmse", 1e-6),
    "max_iteration": ("icp_criteria", "max_iteration", 30),
}


class Icp_parameters():
    def __init__(self, yaml_file='icp_parameters.yaml'):
        here = os.path.dirname(os.path.abspath(__file__))
        with open(os.path.join(here, yaml_file)) as stream:
            tree = yaml.safe_load(stream) or {}
        for attr, (section, key, default) in _FIELDS.items():
            value = (tree.get(section) or {}).get(key, default)
            if value is ...:
                raise KeyError("icp_parameters.yaml: missing %s.%s" % (section, key))
            setattr(self, attr, value)


ICP_PARAMETERS = Icp_parameters()
