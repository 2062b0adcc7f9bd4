"""Import-time singleton of the ICP parameters, same attribute names as the reference's config/config.py:6-34
(max_radius, min_radius, max_height, min_height, voxel_size, radius_gd, max_nn_gd, radius_normals, max_nn,
distance_threshold), read from icp_parameters.yaml next to this module.  Extra, optional: the ICP convergence
criteria (Open3D defaults when the section is absent, as in the reference's yaml)."""
import os

import yaml


class Icp_parameters():
    def __init__(self, yaml_file='icp_parameters.yaml'):
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), yaml_file)
        with open(path) as file:
            cfg = yaml.load(file, Loader=yaml.FullLoader)
        by_radius = cfg.get('filter_by_radius')
        by_height = cfg.get('filter_by_height')
        self.max_radius = by_radius.get('max_radius')
        self.min_radius = by_radius.get('min_radius')
        self.max_height = by_height.get('max_height')
        self.min_height = by_height.get('min_height')
        self.voxel_size = cfg.get('down_sample').get('voxel_size')
        self.radius_gd = cfg.get('filter_ground_plane').get('radius_normals')
        self.max_nn_gd = cfg.get('filter_ground_plane').get('maximum_neighbors')
        self.radius_normals = cfg.get('normals').get('radius_normals')
        self.max_nn = cfg.get('normals').get('maximum_neighbors')
        self.distance_threshold = cfg.get('icp').get('distance_threshold')
        crit = cfg.get('icp_criteria') or {}
        self.relative_fitness = crit.get('relative_fitness', 1e-6)
        self.relative_rmse = crit.get('relative_rmse', 1e-6)
        self.max_iteration = crit.get('max_iteration', 30)


ICP_PARAMETERS = Icp_parameters()
