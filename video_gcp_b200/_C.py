"""ctypes binding of libgcpb200.so (the C ABI declared in include/gcpb200.h).

The library is built in-tree (`python -c "import __graft_entry__ as g; g.build()"` or `make -C
video_gcp_b200/csrc`).  There is NO fallback: if the shared library is missing, `load()` raises.
"""
import ctypes as C
import os

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgcpb200.so")


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("max_candidates", C.c_int), ("attach_cost_mdl", C.c_int),
                ("reserved0", C.c_int), ("decoder_slot_chunk", C.c_int), ("model", C.c_int),
                ("hierarchy_levels", C.c_int), ("max_seq_len", C.c_int), ("tied_layers", C.c_int)]


MODEL_TREE, MODEL_SEQUENTIAL, MODEL_TREE_ADAPTIVE = 0, 1, 2


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int), ("shape", C.c_int64 * 4)]


class RolloutIO(C.Structure):
    _fields_ = [("I_0", C.c_void_p), ("I_g", C.c_void_p), ("images_shared", C.c_int), ("z", C.c_void_p), ("z_host", C.c_void_p),
                ("end_ind", C.c_void_p), ("seed", C.c_uint64), ("B", C.c_int),
                ("e_0", C.c_void_p), ("e_g", C.c_void_p), ("seq_len_logits", C.c_void_p),
                ("end_ind_out", C.c_void_p), ("e_df", C.c_void_p), ("mu_df", C.c_void_p),
                ("log_sigma_df", C.c_void_p), ("images_df", C.c_void_p), ("existence", C.c_void_p),
                ("model_enc_seq", C.c_void_p), ("actions", C.c_void_p), ("regressed_state", C.c_void_p),
                ("distances", C.c_void_p), ("pruned_nodes", C.c_void_p), ("pruned_len", C.c_void_p),
                ("prune_threshold", C.c_float),
                ("decode_kept_only", C.c_int), ("l2_goal", C.c_void_p), ("l2_cost", C.c_void_p), ("l2_dense", C.c_int),
                ("l2_final_step_weight", C.c_float), ("tree_kept_only", C.c_int),
                ("sort_sampled_lengths", C.c_int)]


class SeqIO(C.Structure):
    _fields_ = [("I_0", C.c_void_p), ("I_g", C.c_void_p), ("images_shared", C.c_int), ("z", C.c_void_p),
                ("end_ind", C.c_void_p), ("given_end_ind", C.c_void_p), ("seed", C.c_uint64), ("B", C.c_int),
                ("e_0", C.c_void_p), ("e_g", C.c_void_p), ("seq_len_logits", C.c_void_p), ("end_ind_out", C.c_void_p),
                ("encodings", C.c_void_p), ("mu", C.c_void_p), ("log_sigma", C.c_void_p), ("images", C.c_void_p),
                ("model_enc_seq", C.c_void_p), ("actions", C.c_void_p), ("regressed_state", C.c_void_p)]


LOSS_NAMES = ("len_pred", "action_reconst", "cost_estimation", "state_regression", "dense_img_rec", "kl",
              "existence_predictor", "entropy", "total")          # GCPB200_LOSS_* order


class TrainIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("traj_seq", "pad_mask", "end_ind", "I_0", "I_g", "states", "actions", "eps", "inv_t0", "inv_t1",
                 "cost_start", "cost_end", "cost_target")] + [("B", C.c_int)] + [(n, C.c_void_p) for n in
                ("losses", "nll_per_frame", "kl_per_seq", "e_0", "e_g", "enc_traj_seq", "inf_enc_seq", "seq_len_logits",
                 "e_df", "p_mu", "p_log_sigma", "q_mu", "q_log_sigma", "match_timesteps", "images_df", "existence",
                 "model_enc_seq", "regressed_state", "inv_actions", "cost_pred")]


EXPORTS = {
    # name: (restype, argtypes)
    "gcpb200_last_error": (C.c_char_p, []),
    "gcpb200_version": (C.c_char_p, []),
    "gcpb200_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config)]),
    "gcpb200_destroy": (None, [C.c_void_p]),
    "gcpb200_load_weights": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.c_int]),
    "gcpb200_rollout": (C.c_int, [C.c_void_p, C.POINTER(RolloutIO), C.c_void_p]),
    "gcpb200_seq_rollout": (C.c_int, [C.c_void_p, C.POINTER(SeqIO), C.c_void_p]),
    "gcpb200_cost_l2_seq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                      C.c_void_p, C.c_void_p]),
    "gcpb200_gather_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p]),
    "gcpb200_cost_l2_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_float, C.c_void_p, C.c_void_p]),
    "gcpb200_prune_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gcpb200_cost_l2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                  C.c_void_p, C.c_void_p]),
    "gcpb200_cost_learned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p]),
    "gcpb200_cost_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "gcpb200_infer_action": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gcpb200_forward_loss": (C.c_int, [C.c_void_p, C.POINTER(TrainIO), C.c_void_p]),
    "gcpb200_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gcpb200_refit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gcpb200_sample_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_uint64,
                                       C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "gcpb200_sample_noise_ids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p,
                                           C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "gcpb200_cdist_mean": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "gcpb200_soft_dtw_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gcpb200_soft_dtw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gcpb200_dtw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gcpb200_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gcpb200_sq_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "gcpb200_optim_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                     C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_float,
                                     C.c_void_p]),
    "gcpb200_launch_count": (C.c_int64, [C.c_void_p]),
    "gcpb200_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "gcpb200_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


class GcpB200Error(RuntimeError):
    pass


def bind(path):
    """dlopen a build of the library and set the prototypes of every exported symbol."""
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)            # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """The shipped library (one code path, no environment switches).  Raises when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise GcpB200Error(
            "libgcpb200.so not found at %s -- build it first (`make -C video_gcp_b200/csrc` or "
            "`__graft_entry__.build()`); this package has no CPU or PyTorch fallback." % LIB_PATH)
    _LIB = bind(LIB_PATH)
    return _LIB


def check(rc, lib=None):
    if rc != 0:
        raise GcpB200Error((lib if lib is not None else load()).gcpb200_last_error().decode())
