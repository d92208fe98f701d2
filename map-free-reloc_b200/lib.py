"""ctypes binding of libmfr_b200.so. There is NO CPU fallback: if the library is missing or a
CUDA device is absent, compute entry points raise."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmfr_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "mfr_b200.h")

_lib = None

c_f32p = ctypes.c_void_p
c_ll = ctypes.c_longlong
c_int = ctypes.c_int
c_float = ctypes.c_float
c_vp = ctypes.c_void_p

class LoftrOutputs(ctypes.Structure):
    """mfr_loftr_outputs of include/mfr_b200.h."""
    _fields_ = [("m_total", c_vp), ("m_offsets", c_vp), ("b_ids", c_vp), ("i_ids", c_vp),
                ("j_ids", c_vp), ("mconf", c_vp), ("mkpts0_f", c_vp), ("mkpts1_f", c_vp),
                ("mkpts1_c", c_vp), ("expec_f", c_vp), ("capacity", c_int)]


class SpsgOutputs(ctypes.Structure):
    """mfr_spsg_outputs of include/mfr_b200.h."""
    _fields_ = [("counts", c_vp), ("keypoints", c_vp), ("scores", c_vp), ("descriptors", c_vp),
                ("desc_ld", c_int), ("matches0", c_vp), ("matching_scores0", c_vp), ("m_offsets", c_vp),
                ("mkpts0", c_vp), ("mkpts1", c_vp), ("max_keypoints", c_int)]


c_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/mfr_b200.h one to one
SIGNATURES = {
    "mfr_version": (c_int, []),
    "mfr_device_sm_count": (c_int, []),
    "mfr_conv2d_nhwc_tf32": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_ll, c_vp, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_vp, c_vp, c_ll, c_int, c_float,
                                     c_vp, c_ll, c_vp]),
    "mfr_conv2d_nhwc_bf16": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_ll, c_vp, c_ll, c_int, c_int, c_int,
                                     c_int, c_int, c_vp, c_vp, c_ll, c_int, c_float, c_vp, c_ll, c_int, c_vp]),
    "mfr_loftr_create": (c_int, [c_int, c_int, c_int, c_float, c_int, c_float, c_int, c_pp]),
    "mfr_loftr_destroy": (None, [c_vp]),
    "mfr_loftr_num_weights": (c_int, [c_vp]),
    "mfr_loftr_weight_name": (ctypes.c_char_p, [c_vp, c_int]),
    "mfr_loftr_weight_numel": (c_ll, [c_vp, c_int]),
    "mfr_loftr_set_weight": (c_int, [c_vp, c_int, c_vp, c_ll, c_vp]),
    "mfr_loftr_forward": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "mfr_loftr_set_graph": (None, [c_vp, c_int]),
    "mfr_loftr_get_outputs": (c_int, [c_vp, ctypes.POINTER(LoftrOutputs)]),
    "mfr_loftr_debug_tensor": (c_int, [c_vp, ctypes.c_char_p, c_pp, ctypes.POINTER(c_ll)]),
    "mfr_loftr_num_kernels": (c_int, [c_vp]),
    "mfr_loftr_num_launches": (c_int, [c_vp]),
    "mfr_loftr_profile": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "mfr_spsg_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_float, c_pp]),
    "mfr_spsg_create_ex": (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_float, c_int, c_pp]),
    "mfr_spsg_destroy": (None, [c_vp]),
    "mfr_spsg_num_weights": (c_int, [c_vp]),
    "mfr_spsg_weight_name": (ctypes.c_char_p, [c_vp, c_int]),
    "mfr_spsg_weight_numel": (c_ll, [c_vp, c_int]),
    "mfr_spsg_set_weight": (c_int, [c_vp, c_int, c_vp, c_ll, c_vp]),
    "mfr_spsg_forward": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "mfr_spsg_set_graph": (None, [c_vp, c_int]),
    "mfr_spsg_get_outputs": (c_int, [c_vp, ctypes.POINTER(SpsgOutputs)]),
    "mfr_spsg_debug_tensor": (c_int, [c_vp, ctypes.c_char_p, c_pp, ctypes.POINTER(c_ll)]),
    "mfr_spsg_num_launches": (c_int, [c_vp]),
    "mfr_spsg_forward_superglue": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mfr_coarse_match": (c_int, [c_vp, c_int, c_int, c_int, c_float, c_int, c_float, c_int, c_vp, c_vp,
                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mfr_knn2_l2": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mfr_pose_solver_create": (c_int, [c_int, c_int, c_int, c_pp]),
    "mfr_pose_solver_destroy": (None, [c_vp]),
    "mfr_emat_ransac": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_float, c_int,
                                ctypes.c_ulonglong, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mfr_scale_from_depth": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp, c_int, c_int, c_float, c_vp, c_vp, c_vp, c_vp]),
    "mfr_pnp_ransac": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int,
                               c_float, ctypes.c_ulonglong, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mfr_linear_attention": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "mfr_stem_conv7x7": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "mfr_fine_attention_fused": (c_int, [c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, c_ll, c_int, c_vp]),
    "mfr_icp_refine": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_float, c_int, c_float, c_float,
                               c_vp, c_vp, c_vp, c_vp]),
    "mfr_jpeg_info": (c_int, [c_vp, c_ll, c_vp, c_vp, c_vp]),
    "mfr_jpeg_decode": (c_int, [c_vp, c_ll, c_int, c_vp, c_int, c_int, c_vp]),
    "mfr_resize_linear": (c_int, [c_vp, c_int, c_int, c_int, c_ll, c_int, c_vp, c_int, c_int, c_int, c_float, c_int, c_vp]),
    "mfr_pose_metrics": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "mfr_encoder_tail": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_ll, c_vp, c_int, c_vp]),
    "mfr_debug_probe_overhead": (c_float, [c_int, c_int, c_int, c_int]),
    "mfr_debug_mma_probe": (c_float, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "mfr_debug_time_conv": (c_float, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp]),
    "mfr_procrustes_ransac": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int,
                                      c_float, ctypes.c_ulonglong, c_vp, c_vp, c_vp, c_vp]),
    "mfr_host_five_point": (c_int, [c_vp, c_vp, c_vp]),
    "mfr_host_p3p": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "mfr_host_poly_real_roots": (c_int, [c_vp, c_int, c_vp]),
    "mfr_host_decompose_essential": (c_int, [c_vp, c_vp, c_vp, c_vp]),
}


class MfrError(RuntimeError):
    pass


def load():
    """Loads the shared library (building is done by ``build.py`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MfrError(
            f"{LIB_PATH} not found: run `python map-free-reloc_b200/build.py` "
            "(there is no CPU fallback for the CUDA path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        raise MfrError(f"libmfr_b200 call failed rc={rc} {what}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
