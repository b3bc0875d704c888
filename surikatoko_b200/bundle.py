"""ctypes mirror of include/srk/bundle_c_api.h: bundle files, the binary dump of the flat BA problem."""
import ctypes as C

import numpy as np

from .capi import BAProblem, SrkError, load_library


def _lib():
    L = load_library()
    if not getattr(L, "_bundle_bound", False):
        L.srk_bundle_write.argtypes = [C.c_char_p, C.c_void_p]
        L.srk_bundle_read_header.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.srk_bundle_read.argtypes = [C.c_char_p] + [C.c_void_p] * 6
        L._bundle_bound = True
    return L


def _check(rc):
    if rc < 0:
        raise SrkError(rc, load_library().srk_last_error().decode())
    return rc


def write_bundle(path, problem):
    """srk_bundle_write: dump a BAProblem."""
    st = problem.c_struct()
    _check(_lib().srk_bundle_write(str(path).encode(), C.addressof(st)))


def read_bundle_header(path):
    nc, npnt, no, sk, f0 = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_double()
    _check(_lib().srk_bundle_read_header(str(path).encode(), C.byref(nc), C.byref(npnt), C.byref(no), C.byref(sk), C.byref(f0)))
    return dict(n_cams=nc.value, n_points=npnt.value, n_obs=no.value, shared_K=bool(sk.value), f0=f0.value)


def read_bundle(path):
    """srk_bundle_read: -> BAProblem (verifies magic, sizes, length and checksum)."""
    h = read_bundle_header(path)
    obs_cam = np.zeros(h["n_obs"], dtype=np.int32); obs_point = np.zeros(h["n_obs"], dtype=np.int32)
    obs_xy = np.zeros((h["n_obs"], 2)); points = np.zeros((h["n_points"], 3)); cams = np.zeros((h["n_cams"], 12))
    K = np.zeros((1 if h["shared_K"] else h["n_cams"], 9))
    _check(_lib().srk_bundle_read(str(path).encode(), obs_cam.ctypes.data, obs_point.ctypes.data, obs_xy.ctypes.data, points.ctypes.data, cams.ctypes.data,
                                  K.ctypes.data))
    return BAProblem(obs_cam, obs_point, obs_xy, points, cams, K, h["shared_K"], h["f0"])


def _main(argv):
    """python -m surikatoko_b200.bundle write <M> <N> <obs_per_point> <path> [seed]   |   python -m surikatoko_b200.bundle info <path>"""
    if len(argv) >= 5 and argv[0] == "write":
        from . import scenes
        pr = scenes.ring_scene(int(argv[1]), int(argv[2]), int(argv[3]), seed=int(argv[5]) if len(argv) > 5 else 1234)
        write_bundle(argv[4], pr)
        print("wrote", argv[4], read_bundle_header(argv[4]))
        return 0
    if len(argv) == 2 and argv[0] == "info":
        print(read_bundle_header(argv[1]))
        return 0
    print(_main.__doc__)
    return 2


if __name__ == "__main__":
    import sys
    sys.exit(_main(sys.argv[1:]))
