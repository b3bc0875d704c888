"""ctypes mirror of include/srk/frontend_c_api.h: the input front end of the reference's two BA demos
(demo-bundle-adj-dinosaur.cpp:79-187, demo-bundle-adj-circle-grid.cpp) -- projection-matrix decomposition, text matrix
reader, batched triangulation on the GPU.  Names follow the reference (obs-geom.h, mat-serialization.h)."""
import ctypes as C

import numpy as np

from .capi import SrkError, load_library


def _lib():
    L = load_library()
    if not getattr(L, "_frontend_bound", False):
        L.srk_triangulate_tracks.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.srk_decompose_proj_mat.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.srk_read_matrix_from_file.argtypes = [C.c_char_p, C.c_char, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.srk_triangulate_last_kernel_ms.restype = C.c_double
        L._frontend_bound = True
    return L


def _check(rc):
    if rc < 0:
        raise SrkError(rc, load_library().srk_last_error().decode())
    return rc


def Triangulate3DPointByLeastSquares(track_begin, obs_frame, obs_xy, proj_mats, f0, device=0, out=None):
    """Batched obs-geom.cpp:679-727.  proj_mats: [n_frames, 3, 4] (f0-scaled projection matrices); returns [n_tracks, 3]
    (written into `out` when given, e.g. a pinned buffer)."""
    tb = np.ascontiguousarray(track_begin, dtype=np.int64)
    fr = np.ascontiguousarray(obs_frame, dtype=np.int32)
    xy = np.ascontiguousarray(obs_xy, dtype=np.float64).reshape(-1, 2)
    P = np.asarray(proj_mats, dtype=np.float64).reshape(-1, 3, 4)
    pm = np.ascontiguousarray(P.transpose(0, 2, 1)).reshape(-1, 12)      # column-major per frame
    if out is None:
        out = np.zeros((len(tb) - 1, 3))
    assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == 3 * (len(tb) - 1)
    _check(_lib().srk_triangulate_tracks(device, len(tb) - 1, len(fr), P.shape[0], tb.ctypes.data, fr.ctypes.data, xy.ctypes.data, pm.ctypes.data,
                                         float(f0), out.ctypes.data))
    return out


def last_triangulation_kernel_ms():
    return float(_lib().srk_triangulate_last_kernel_ms())


def DecomposeProjMat(proj_mat):
    """obs-geom.cpp:606-677: P[3x4] -> (ok, scale_factor, K[3x3], (R, T) of the direct camera pose), P = scale K R^T [I | -T]."""
    P = np.ascontiguousarray(np.asarray(proj_mat, dtype=np.float64).reshape(3, 4).T).reshape(12)
    scale = C.c_double()
    K = np.zeros(9); pose = np.zeros(12)
    rc = _check(_lib().srk_decompose_proj_mat(P.ctypes.data, C.byref(scale), K.ctypes.data, pose.ctypes.data))
    if rc != 0:
        return False, 0.0, None, None
    return True, scale.value, K.reshape(3, 3).T.copy(), (pose[3:].reshape(3, 3).T.copy(), pose[:3].copy())


def ReadMatrixFromFile(path, delimiter):
    """mat-serialization.cpp:12-87: row-major text matrix -> ndarray [rows, cols]; raises SrkError with the reference's message."""
    rows, cols = C.c_int64(), C.c_int64()
    L = _lib()
    d = delimiter.encode()[:1]
    _check(L.srk_read_matrix_from_file(str(path).encode(), d, None, 0, C.byref(rows), C.byref(cols)))
    data = np.zeros(max(1, rows.value * cols.value))
    _check(L.srk_read_matrix_from_file(str(path).encode(), d, data.ctypes.data, data.size, C.byref(rows), C.byref(cols)))
    return data[:rows.value * cols.value].reshape(rows.value, cols.value)
