"""``gravomg.util`` of the reference, served by gravo_mg_b200."""
from gravo_mg_b200.util import *  # noqa: F401,F403
from gravo_mg_b200.util import (coalesce_edges, face_area, homogenize_edges, knn, knn_undirected,  # noqa: F401
                                neighbors_from_faces, neighbors_from_stiffness, normalize_area,
                                normalize_axes, normalize_bounding_box)
