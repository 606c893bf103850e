"""ctypes binding of the C ABI in include/gdss_b200.h (libgdss_b200.so, sm_100a CUDA).

There is no CPU fallback: importing works on a CPU-only host (so that the host-side logic and
the symbol table can be tested), but every compute entry point fails loudly without a CUDA device,
and a missing shared library is an ImportError-grade failure, never a silent downgrade.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GDSS_PRECISION=tf32 (read once, at the first call) selects the reduced-precision build of the same sources: single-pass
# TF32 contractions instead of the 3xTF32 split, raw scores within 3e-2 rel-L2 of the reference instead of 1e-4.  Anything
# else (the default) is the fp32-parity library.
PHASE_CORRECTOR, PHASE_PREDICTOR = 1, 2          # include/gdss_b200.h: GDSS_PHASE_*
PRECISION = "tf32" if os.environ.get("GDSS_PRECISION", "").lower() == "tf32" else "fp32"
LIB_PATH = os.path.join(_HERE, "csrc", "libgdss_b200_tf32.so" if PRECISION == "tf32" else "libgdss_b200.so")
if os.environ.get("GDSS_LIB"):          # development only: A/B of two builds of the same ABI (tools/gpu_ab.sh)
    LIB_PATH = os.environ["GDSS_LIB"]

NET_X, NET_X_GMH, NET_A = 0, 1, 2
SDE_VP, SDE_VE, SDE_SUBVP = 0, 1, 2
PRED_EULER, PRED_REVERSE, PRED_S4 = 0, 1, 2
CORR_NONE, CORR_LANGEVIN = 0, 1


class NetDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("model_type", "max_feat_num", "max_node_num", "nhid", "depth", "num_linears",
                                         "c_init", "c_hid", "c_final", "adim", "num_heads", "reserved")]


class SdeDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("num_scales", C.c_int32), ("lo", C.c_double), ("hi", C.c_double)]


class SamplerDesc(C.Structure):
    _fields_ = [("sde_x", SdeDesc), ("sde_adj", SdeDesc), ("predictor", C.c_int32), ("corrector", C.c_int32),
                ("snr", C.c_double), ("scale_eps", C.c_double), ("eps", C.c_double), ("n_steps", C.c_int32),
                ("probability_flow", C.c_int32), ("denoise", C.c_int32), ("reserved", C.c_int32)]


STEP_COEF_FLOATS = 24      # sizeof(gdss_step_coef) / 4
# column index of each field of gdss_step_coef (x, adj pairs)
COL = dict(cs=0, alpha=2, pa=4, pb=6, pc=8, mu1=10, sig1=12, drift=14, mu2=16, sig2=18, snr=20, seps=21, t=22)


class GdssError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once.  Raises if it has not been built (python -c 'import
    __graft_entry__ as g; g.build()' or make -C gdss_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise GdssError(f"{LIB_PATH} is missing: build it with `make -C gdss_b200/csrc` "
                        "(there is no CPU or PyTorch fallback for the GDSS kernels)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, u64, fp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p
    sig = {
        "gdss_abi_version": (C.c_int, []),
        "gdss_device_count": (C.c_int, []),
        "gdss_error_string": (C.c_char_p, [C.c_int]),
        "gdss_packed_floats": (i64, [C.POINTER(NetDesc), i32]),
        "gdss_num_state_tensors": (i32, [C.POINTER(NetDesc)]),
        "gdss_pack_weights": (C.c_int, [C.POINTER(NetDesc), i32, C.POINTER(vp), C.POINTER(i64), i32, fp, i64]),
        "gdss_analyze_weights": (C.c_int, [C.POINTER(NetDesc), i32, fp, C.POINTER(i32), i32]),
        "gdss_ctx_create": (vp, [i32, i32, C.POINTER(NetDesc), fp, C.POINTER(NetDesc), fp]),
        "gdss_ctx_destroy": (None, [vp]),
        "gdss_workspace_bytes": (i64, [vp, i32]),
        "gdss_workspace_offset": (i64, [vp, i32, C.c_char_p]),
        "gdss_ctx_query": (i64, [vp, C.c_char_p]),
        "gdss_score_forward": (C.c_int, [vp, fp, fp, fp, fp, fp, i32, i32, vp, i64, vp]),
        "gdss_sample": (C.c_int, [vp, C.POINTER(SamplerDesc), fp, i32, i32, fp, fp, fp, fp, fp, i32, i32, u64, i64, i64,
                                   fp, fp, fp, vp, i32, vp, i64, vp]),
        "gdss_solver_phase": (C.c_int, [vp, C.POINTER(SamplerDesc), fp, i32, i32, fp, fp, fp, fp, fp, i32, u64, i64, i64, fp, fp, fp,
                                         vp, vp, i64, vp]),
        "gdss_dsm_scratch_floats": (i64, [vp, i32]),
        "gdss_dsm_loss": (C.c_int, [vp, fp, fp, fp, fp, fp, i32, i32, fp, i64, fp, vp, i64, vp]),
        "gdss_train_workspace_bytes": (i64, [vp, i32]),
        "gdss_dsm_loss_backward": (C.c_int, [vp, fp, fp, fp, fp, fp, i32, i32, fp, fp, fp, vp, i64, vp]),
        "gdss_optim_scratch_floats": (i64, []),
        "gdss_adam_ema_step": (C.c_int, [fp, fp, fp, fp, fp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                          C.c_float, i32, C.c_float, i64, fp, fp, fp, vp]),
        "gdss_adam_step_consts": (C.c_int, [C.c_float, C.c_float, C.c_float, i32, C.c_float, i64, fp]),
        "gdss_adam_ema_step_dev": (C.c_int, [fp, fp, fp, fp, fp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                              fp, fp, fp, fp, vp]),
        "gdss_nccl_allreduce": (C.c_int, [vp, fp, i64, vp]),
        "gdss_nccl_unique_id": (C.c_int, [vp]),
        "gdss_nccl_comm_create": (C.c_int, [vp, i32, i32, C.POINTER(vp)]),
        "gdss_nccl_comm_destroy": (C.c_int, [vp]),
        "gdss_nccl_allgather": (C.c_int, [vp, fp, fp, i64, vp]),
        "gdss_mask_x": (C.c_int, [fp, fp, i32, i32, i32, vp]),
        "gdss_mask_adjs": (C.c_int, [fp, fp, i32, i32, i32, vp]),
        "gdss_gen_noise": (C.c_int, [fp, fp, fp, i32, i32, i32, i32, u64, u64, i64, vp]),
        "gdss_pow_tensor": (C.c_int, [fp, fp, i32, i32, i32, vp]),
        "gdss_quantize": (C.c_int, [fp, fp, C.c_float, i64, vp]),
        "gdss_quantize_mol": (C.c_int, [fp, vp, i64, vp]),
        "gdss_decode_mol": (C.c_int, [fp, fp, vp, vp, i32, i32, i32, vp]),
        "gdss_init_flags": (C.c_int, [fp, fp, i32, i32, u64, i64, vp]),
        "gdss_launch_count": (i64, []),
        "gdss_prof_begin": (C.c_int, [vp]),
        "gdss_prof_end": (C.c_int, [fp, vp, i32]),
        "gdss_prof_name": (C.c_char_p, [i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    L._gdss_symbols = tuple(sig)
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().gdss_error_string(int(rc))
        raise GdssError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")


def require_cuda():
    if lib().gdss_device_count() <= 0:
        raise GdssError("no CUDA device visible: the GDSS kernels are sm_100a CUDA only and have no CPU fallback")
