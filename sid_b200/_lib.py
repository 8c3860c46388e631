"""ctypes binding of include/sidgpu.h.  Loading fails loudly when libsidgpu.so is missing:
there is no Python or CPU implementation of the path to fall back to."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsidgpu.so")

c_u64 = ctypes.c_uint64
c_u64_p = ctypes.POINTER(ctypes.c_uint64)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)
c_double_p = ctypes.POINTER(ctypes.c_double)


class Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("max_chunk_bytes", ctypes.c_size_t), ("max_sites", ctypes.c_size_t),
                ("table_log2", ctypes.c_int), ("stream", ctypes.c_void_p)]


class Params(ctypes.Structure):
    _fields_ = [("method", ctypes.c_int), ("estimate_prior", ctypes.c_int), ("prior", ctypes.c_double),
                ("error_threshold", ctypes.c_double), ("significance_level", ctypes.c_double),
                ("fit_given", ctypes.c_int), ("fit_pi", ctypes.c_double), ("fit_eps", ctypes.c_double),
                ("fit_nd", ctypes.c_double * 4), ("het_only", ctypes.c_int), ("want_strands", ctypes.c_int)]


class Columns(ctypes.Structure):
    _fields_ = [("d_pos", ctypes.c_void_p), ("d_name_ref", ctypes.c_void_p), ("d_label", ctypes.c_void_p), ("d_gt", ctypes.c_void_p),
                ("d_hom_conf", ctypes.c_void_p), ("d_het_conf", ctypes.c_void_p), ("d_profile", ctypes.c_void_p), ("d_fwd", ctypes.c_void_p)]


class SitesView(ctypes.Structure):
    _fields_ = [("n_sites", c_u64), ("d_profile", ctypes.c_void_p), ("d_pos", ctypes.c_void_p),
                ("d_slot", ctypes.c_void_p), ("d_line_off", ctypes.c_void_p), ("d_name_ref", ctypes.c_void_p),
                ("d_names", ctypes.c_void_p), ("names_bytes", c_u64), ("d_fwd", ctypes.c_void_p)]


class UniqueView(ctypes.Structure):
    _fields_ = [("n_unique", c_u64), ("d_profile", ctypes.c_void_p), ("d_count", ctypes.c_void_p),
                ("nd", ctypes.c_double * 4), ("nd_sums", c_u64 * 5)]


class Fit(ctypes.Structure):
    _fields_ = [("pi", ctypes.c_double), ("eps", ctypes.c_double), ("fval", ctypes.c_double),
                ("iterations", ctypes.c_int), ("evaluations", ctypes.c_int), ("converged", ctypes.c_int)]


# name -> (restype, argtypes); every symbol include/sidgpu.h declares
IO_READ = ctypes.CFUNCTYPE(ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
IO_WRITE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
IO_REWIND = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)


class BgzfBlock(ctypes.Structure):
    _fields_ = [("c_off", ctypes.c_uint64), ("out_off", ctypes.c_uint64), ("c_len", ctypes.c_uint32), ("isize", ctypes.c_uint32),
                ("crc", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class Io(ctypes.Structure):                  # sidgpu_io
    _fields_ = [("read", IO_READ), ("write", IO_WRITE), ("rewind", IO_REWIND), ("user", ctypes.c_void_p)]


PROTOTYPES = {
    "sidgpu_create": (ctypes.c_int, [ctypes.POINTER(Config), c_void_pp]),
    "sidgpu_destroy": (None, [ctypes.c_void_p]),
    "sidgpu_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "sidgpu_version": (ctypes.c_char_p, []),
    "sidgpu_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "sidgpu_malloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, c_void_pp]),
    "sidgpu_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "sidgpu_malloc_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, c_void_pp]),
    "sidgpu_free_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "sidgpu_memcpy_h2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "sidgpu_memcpy_d2h": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "sidgpu_memcpy_d2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "sidgpu_memcpy_d2d_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "sidgpu_tokenize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                       ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(SitesView)]),
    "sidgpu_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Params)]),
    "sidgpu_feed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                   ctypes.c_size_t, c_u64_p]),
    "sidgpu_feed_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_size_t, c_u64_p, c_u64_p, c_u64_p]),
    "sidgpu_finish": (ctypes.c_int, [ctypes.c_void_p]),
    "sidgpu_emit_csv": (ctypes.c_int, [ctypes.c_void_p, c_u64, c_u64, ctypes.c_void_p, ctypes.c_size_t, c_u64_p, c_u64_p]),
    "sidgpu_feed_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(c_u64)]),
    "sidgpu_emit_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]),
    "sidgpu_stream_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                          ctypes.POINTER(c_u64), ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]),
    "sidgpu_emit_columns": (ctypes.c_int, [ctypes.c_void_p, c_u64, c_u64, ctypes.POINTER(Columns)]),
    "sidgpu_names": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, ctypes.POINTER(c_u64)]),
    "sidgpu_emit_records": (ctypes.c_int, [ctypes.c_void_p, c_u64, c_u64, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "sidgpu_call_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_size_t, c_u64_p, c_u64_p, c_u64_p]),
    "sidgpu_call_io": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.POINTER(Io), c_u64_p, c_u64_p, c_u64_p]),
    "sidgpu_call_io_bgzf": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.POINTER(Io), c_u64_p, c_u64_p, c_u64_p]),
    "sidgpu_call_host_bgzf": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                             c_u64_p, c_u64_p, c_u64_p]),
    "sidgpu_bgzf_scan": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "sidgpu_inflate_bgzf": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]),
    "sidgpu_histogram": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(UniqueView)]),
    "sidgpu_count_unique": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64, ctypes.c_uint32, ctypes.POINTER(UniqueView)]),
    "sidgpu_count_unique_weighted": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, c_u64, ctypes.c_uint32,
                                                    ctypes.POINTER(UniqueView)]),
    "sidgpu_set_fit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, c_double_p]),
    "sidgpu_set_global_histogram": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, c_u64]),
    "sidgpu_lynch_objective_partial": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_double, ctypes.c_double,
                                                      ctypes.c_void_p]),
    "sidgpu_lynch_objective": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_double, ctypes.c_double, c_double_p]),
    "sidgpu_lynch_fit": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.POINTER(Fit)]),
    "sidgpu_finish_global": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64]),
    "sidgpu_session_fit": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Fit), c_double_p, c_u64_p]),
    "sidgpu_bh_adjust": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64, ctypes.c_void_p]),
    "sidgpu_lr_test": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, c_u64, ctypes.c_void_p]),
    "sidgpu_profile_loglik": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64, c_double_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "sidgpu_qualities": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, c_u64_p]),
    "sidgpu_strand_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, c_u64, ctypes.c_void_p, ctypes.c_void_p]),
    "sidgpu_read_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, c_u64, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sidgpu_read_fill": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, c_u64, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "sidgpu_format_g": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_u64, ctypes.c_void_p]),
    "sidgpu_launch_count": (c_u64, [ctypes.c_void_p]),
    "sidgpu_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "sidgpu_kernel_times": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_u64_p]),
}

_lib = None


def load():
    """Returns the loaded library; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("sid_b200: %s is missing -- run `python -m sid_b200.build` (nvcc, sm_100a). "
                              "There is no CPU fallback for this path." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
