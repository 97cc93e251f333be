"""Generates tests/golden/util_reference.npz by importing the REFERENCE's own util.py.

Run in the build container only (``/root/reference`` does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference's compiled solver cannot be built offline (needs libigl + Eigen from the
network), so the only reference code that can run here is its pure-Python input preparation,
gravomg_bindings/src/gravomg/util.py. Its outputs on small synthetic inputs are stored as the
golden vectors that pin gravo_mg_b200/util.py (tests/test_util_golden.py).
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_UTIL = "/root/reference/gravomg_bindings/src/gravomg/util.py"


def main():
    spec = importlib.util.spec_from_file_location("reference_util", REF_UTIL)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    from gravo_mg_b200 import synth  # only the mesh generators (inputs), not the code under test

    out = {}
    V, F = synth.icosphere(2)        # 162 vertices
    Vt, Ft = synth.torus_grid(12, 9)  # 108 vertices, valence 6
    rng = np.random.default_rng(7)
    P = rng.standard_normal((200, 3))
    for name, (v, f) in {"ico": (V, F), "torus": (Vt, Ft)}.items():
        S = synth.cotangent_stiffness(v, f, fmt="csc")  # igl.cotmatrix returns CSC
        out[f"{name}_V"] = v
        out[f"{name}_F"] = f
        out[f"{name}_neigh_stiffness"] = ref.neighbors_from_stiffness(S)
        out[f"{name}_neigh_faces"] = ref.neighbors_from_faces(f)
        out[f"{name}_face_area"] = ref.face_area(v, f)
        out[f"{name}_normalize_area"] = ref.normalize_area(v, f)
    out["cloud_P"] = P
    out["cloud_knn6"] = ref.knn(P, 6)
    out["cloud_knn_undirected6"] = ref.knn_undirected(P, 6)
    out["cloud_normalize_bbox"] = ref.normalize_bounding_box(P)
    out["cloud_normalize_axes"] = ref.normalize_axes(P)
    ei = np.array([3, 0, 2, 2, 1, 0, 3, 3, 1, 0])
    ej = np.array([1, 2, 0, 3, 3, 1, 0, 2, 0, 3])
    ci, cj = ref.coalesce_edges(ei, ej)
    out["edges_i"], out["edges_j"] = ei, ej
    out["coalesce_i"], out["coalesce_j"] = ci, cj
    out["homogenize"] = ref.homogenize_edges(ci, cj)
    path = os.path.join(HERE, "util_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
