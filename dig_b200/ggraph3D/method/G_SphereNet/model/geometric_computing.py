"""G-SphereNet's private geometry with the reference's call signatures, on the sm_100a kernels.

reference: dig/ggraph3D/method/G_SphereNet/model/geometric_computing.py
           :22-51  xyztoda(pos, edge_index, num_nodes)            -> dist, angle, i, j, idx_kj, idx_ji
           :54-104 xyztodat(pos, edge_index, num_nodes, batch)    -> dist, angle, torsion, i, j, idx_kj, idx_ji

Only the geometry of the generator's SphereNet copy lives here: distances, triplets and angles are
those of dig.threedgraph's xyz_to_dat; the torsion of a triplet (k -> j -> i) uses ONE reference atom -- the nearest
neighbour of j inside its graph (torch_cluster.knn_graph semantics), or the second nearest when the nearest is i --
instead of the minimum over all neighbours, and is mapped to (0, 2 pi].
"""
from .....threedgraph.utils.geometric_computing import xyz_to_dat


def xyztoda(pos, edge_index, num_nodes):
    return xyz_to_dat(pos, edge_index, num_nodes, use_torsion=False)


def xyztodat(pos, edge_index, num_nodes, batch):
    if batch.shape != (pos.size(0),):
        raise ValueError("batch must be [N]")
    return xyz_to_dat(pos, edge_index, num_nodes, use_torsion=True, _knn_batch=batch.contiguous())
