from .geometric_computing import radius_graph, xyz_to_dat

__all__ = ['xyz_to_dat', 'radius_graph']
