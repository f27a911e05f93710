"""ctypes loader for libargus_b200.so (the C ABI of include/argus_b200.h).

There is no CPU fallback: if the extension is missing this raises, and every compute entry point returns
ARGUS_E_CUDA (-2) when no CUDA device is present, which the wrappers turn into ``ArgusCudaError``.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libargus_b200.so")

ARGUS_OK, ARGUS_E_SBA, ARGUS_E_CUDA, ARGUS_E_ARG, ARGUS_E_COMM = 0, -1, -2, -3, -4
COMPAT_SBA16, PROFILE = 1, 2

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p)


class ArgusCudaError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); also the list of symbols include/argus_b200.h declares (tests check every one)
SIGNATURES = {
    "argus_b200_version": (ctypes.c_char_p, []),
    "argus_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, c_double_p, c_double_p]),
    "argus_ba_motstr_kdrts": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "argus_ba_mot_kdrts": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "argus_ba_str_kdrts": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "argus_pair_triangulate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "argus_dlt_fit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p]),
    "argus_ba_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p]),
    "argus_ba_destroy": (None, [ctypes.c_void_p]),
    "argus_ba_set_comm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ALLREDUCE_FN, ctypes.c_void_p]),
    "argus_set_default_device": (ctypes.c_int, [ctypes.c_int]),
    "argus_nccl_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "argus_ba_comm_init_nccl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "argus_ba_set_comm_nccl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "argus_ba_get_comm_nccl": (ctypes.c_void_p, [ctypes.c_void_p]),
    "argus_ba_set_graph": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "argus_ba_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "argus_ba_set_mode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "argus_ba_set_camera_fix": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "argus_ba_set_wand": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]),
    "argus_ba_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "argus_ba_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "argus_ba_download": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "argus_ba_num_observations": (ctypes.c_int64, [ctypes.c_void_p]),
    "argus_ba_get_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), c_double_p,
                                            ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "argus_ba_get_counters": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "argus_ba_project": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "argus_ba_normal_equations": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double] + [ctypes.c_void_p] * 8),
    "argus_dlt_triangulate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "argus_dlt_reproj_error": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "argus_dlt_reconstruct": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "argus_dlt_bootstrap": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]),
    "argus_undistort_points": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "argus_dlt_triangulate_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "argus_dlt_reconstruct_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "argus_dlt_reproj_error_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ArgusCudaError(
            "argus_gui_b200: %s is missing - build it with `python -m argus_gui_b200.build` "
            "(there is no CPU fallback for the BA / DLT path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == ARGUS_E_CUDA:
        raise ArgusCudaError("%s: CUDA error or no CUDA device (argus_gui_b200 has no CPU fallback)" % what)
    if rc == ARGUS_E_ARG:
        raise ValueError("%s: invalid argument" % what)
    if rc == ARGUS_E_COMM:
        raise RuntimeError("%s: the all-reduce (NCCL or hook) failed" % what)
    return rc
