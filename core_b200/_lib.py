"""ctypes loader of the in-tree C-ABI library core_b200/lib/libmag.so (include/mag.h).

The product has no CPU fallback: if the library is missing or no CUDA device is present,
loading / mag_create fails loudly.  Nothing under oracle/ is ever imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmag.so")

# every symbol include/mag.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "mag_create", "mag_destroy", "mag_last_error", "mag_set_stream", "mag_synchronize",
    "mag_set_mesh", "mag_set_mesh_2d", "mag_set_coords",
    "mag_set_metric_identity", "mag_set_metric_uniform_refiner", "mag_set_metric_iso", "mag_set_metric_aniso", "mag_set_metric_logm",
    "mag_set_flags", "mag_clear_flag", "mag_reset_layer", "mag_sweep", "mag_sweep_host", "mag_resweep_host", "mag_set_mark_bytes", "mag_get_mark_bytes", "mag_element_weights", "mag_prism_weights", "mag_split_vertices", "mag_cavity_quality", "mag_collapse_quality", "mag_short_edge_test",
    "mag_sliver_codes", "mag_linear_qualities",
    "mag_get_edge_lengths", "mag_get_qualities", "mag_get_flags", "mag_get_layer_ok", "mag_get_stats",
    "mag_get_near_threshold", "mag_set_metric_logm_from_frames",
    "mag_timing_begin", "mag_timing_read", "mag_launch_count", "mag_get_row_layout",
    "mag_comm_unique_id", "mag_comm_init", "mag_set_edge_links", "mag_reconcile_edge_flags",
    "mag_smb_read", "mag_smb_free", "mag_smb_last_error", "mag_smb_get", "mag_smb_vertex_field", "mag_set_mesh_smb",
    "mag_sync_edge_flags", "mag_sweep_reconciled", "mag_check_edge_flag_consistency", "mag_allreduce_stats",
]


class MagStats(C.Structure):
    _fields_ = [("n_split", C.c_int64), ("n_collapse", C.c_int64), ("n_bad", C.c_int64),
                ("n_edges_evaluated", C.c_int64), ("n_elems_evaluated", C.c_int64),
                ("n_near_threshold", C.c_int64), ("n_layer_unsafe", C.c_int64),
                ("n_flag_mismatch", C.c_int64),
                ("min_quality", C.c_double), ("max_length", C.c_double), ("sum_length", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MagHostPart(C.Structure):
    """mag_host_part of include/mag.h"""
    _fields_ = [("nv", C.c_int64), ("xyz", C.c_void_p), ("ne", C.c_int64), ("edge_v", C.c_void_p),
                ("nt", C.c_int64), ("tet_v", C.c_void_p), ("edge_owned", C.c_void_p), ("elem_owned", C.c_void_p),
                ("kind", C.c_int), ("field_a", C.c_void_p), ("field_b", C.c_void_p),
                ("edge_flags", C.c_void_p), ("elem_flags", C.c_void_p), ("slice_entities", C.c_int64)]


class MagHostUpdate(C.Structure):
    """mag_host_update of include/mag.h"""
    _fields_ = [("xyz", C.c_void_p), ("kind", C.c_int), ("field_a", C.c_void_p), ("field_b", C.c_void_p),
                ("edge_marks", C.c_void_p), ("elem_marks", C.c_void_p)]


class MagHostMarks(C.Structure):
    _fields_ = [("edge_marks", C.c_void_p), ("elem_marks", C.c_void_p), ("edge_lengths", C.c_void_p), ("qualities", C.c_void_p)]


class MagSmbArrays(C.Structure):
    """mag_smb_arrays of include/mag.h"""
    _fields_ = [("dim", C.c_int), ("version", C.c_int), ("nparts", C.c_int)] + \
               [(k, C.c_int64) for k in ("nv", "ne", "ntri", "nquad", "nt", "np", "npy", "nhex")] + \
               [(k, C.c_void_p) for k in ("xyz", "edge_v", "tri_v", "tet_v", "prism_v", "pyr_v")]


class MagHostResult(C.Structure):
    _fields_ = [("edge_lengths", C.c_void_p), ("qualities", C.c_void_p), ("edge_flags", C.c_void_p),
                ("elem_flags", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "core_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u32, f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_double
    L.mag_create.argtypes = [C.POINTER(vp), C.c_int]
    L.mag_destroy.argtypes = [vp]
    L.mag_destroy.restype = None
    L.mag_last_error.argtypes = [vp]
    L.mag_last_error.restype = C.c_char_p
    L.mag_set_stream.argtypes = [vp, vp]
    L.mag_synchronize.argtypes = [vp]
    L.mag_set_mesh.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp]
    L.mag_set_mesh_2d.argtypes = [vp, i64, vp, i64, vp, i64, vp, vp, vp]
    L.mag_set_coords.argtypes = [vp, vp]
    L.mag_set_metric_identity.argtypes = [vp]
    L.mag_set_metric_uniform_refiner.argtypes = [vp]
    L.mag_set_metric_iso.argtypes = [vp, vp]
    L.mag_set_metric_aniso.argtypes = [vp, vp, vp]
    L.mag_set_metric_logm.argtypes = [vp, vp]
    L.mag_set_flags.argtypes = [vp, vp, vp]
    L.mag_clear_flag.argtypes = [vp, C.c_int, i32]
    L.mag_reset_layer.argtypes = [vp, vp, C.POINTER(i64)]
    L.mag_sweep.argtypes = [vp, u32, f64, f64, f64, C.c_int, C.c_int]
    L.mag_sweep_reconciled.argtypes = [vp, u32, f64, f64, f64, C.c_int, C.c_int, i32]
    L.mag_sweep_host.argtypes = [vp, C.POINTER(MagHostPart), C.POINTER(MagHostResult), u32, f64, f64, f64, C.c_int, C.c_int,
                                 C.POINTER(MagStats)]
    L.mag_resweep_host.argtypes = [vp, C.POINTER(MagHostUpdate), C.POINTER(MagHostMarks), u32, f64, f64, f64, C.c_int, C.c_int,
                                   C.POINTER(MagStats)]
    L.mag_set_mark_bytes.argtypes = [vp, vp, vp]
    L.mag_get_mark_bytes.argtypes = [vp, vp, vp]
    L.mag_element_weights.argtypes = [vp, f64, f64, C.c_int, vp]
    L.mag_collapse_quality.argtypes = [vp, i64, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.mag_prism_weights.argtypes = [vp, vp, f64, f64, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.mag_cavity_quality.argtypes = [vp, i64, vp, vp, C.c_int, C.c_int, vp, vp]
    L.mag_short_edge_test.argtypes = [vp, vp, f64, vp, C.POINTER(i64), C.POINTER(i64)]
    L.mag_sliver_codes.argtypes = [vp, vp, f64, C.c_int, vp, vp]
    L.mag_linear_qualities.argtypes = [C.c_int, i64, vp, vp, vp, C.POINTER(i64)]
    L.mag_split_vertices.argtypes = [vp, C.c_int, i64, C.POINTER(i64), vp, vp, vp, vp]
    L.mag_get_edge_lengths.argtypes = [vp, vp]
    L.mag_get_qualities.argtypes = [vp, vp]
    L.mag_get_flags.argtypes = [vp, vp, vp]
    L.mag_get_layer_ok.argtypes = [vp, vp, vp]
    L.mag_get_stats.argtypes = [vp, C.POINTER(MagStats)]
    L.mag_get_near_threshold.argtypes = [vp, C.c_int, vp, i64, C.POINTER(i64)]
    L.mag_set_metric_logm_from_frames.argtypes = [vp, vp, vp, C.c_int, vp]
    L.mag_timing_begin.argtypes = [vp, C.c_int]
    L.mag_timing_read.argtypes = [vp, vp, C.POINTER(C.c_int)]
    L.mag_launch_count.argtypes = [vp]
    L.mag_launch_count.restype = i64
    L.mag_get_row_layout.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.mag_comm_unique_id.argtypes = [vp]
    L.mag_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.mag_set_edge_links.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.mag_reconcile_edge_flags.argtypes = [vp, i32]
    L.mag_sync_edge_flags.argtypes = [vp, i32]
    L.mag_smb_read.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.mag_smb_free.argtypes = [vp]
    L.mag_smb_free.restype = None
    L.mag_smb_last_error.argtypes = [vp]
    L.mag_smb_last_error.restype = C.c_char_p
    L.mag_smb_get.argtypes = [vp, C.POINTER(MagSmbArrays)]
    L.mag_smb_vertex_field.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(vp)]
    L.mag_set_mesh_smb.argtypes = [vp, vp]
    L.mag_check_edge_flag_consistency.argtypes = [vp, i32, C.POINTER(i64)]
    L.mag_allreduce_stats.argtypes = [vp, C.POINTER(MagStats)]
    _lib = L
    return L
