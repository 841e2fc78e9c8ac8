"""ctypes declarations mirroring ``include/smcdet_b200.h`` (structs, enums, prototypes)."""

import ctypes as C

ABI_VERSION = 8

MODEL_GAUSS_POISSON, MODEL_M71_NORMAL = 0, 1
COUNT_DISCRETE_UNIFORM, COUNT_POISSON, COUNT_NONE = 0, 1, 2
FLUX_PARETO, FLUX_TRUNCATED_PARETO, FLUX_NORMAL = 0, 1, 2
RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC = 0, 1
STATUS_OUT_OF_BOX = 1
STATUS_BAD_TAPE = 2

E_INVALID, E_UNSUPPORTED, E_TOO_LARGE = -1, -2, -3


class ModelParams(C.Structure):
    _fields_ = [
        ("model_kind", C.c_int32), ("psf_radius", C.c_int32),
        ("psf_stdev", C.c_float),
        ("sigma1", C.c_float), ("sigma2", C.c_float), ("sigmap", C.c_float),
        ("beta", C.c_float), ("b", C.c_float), ("p0", C.c_float),
        ("psf_norm", C.c_float),
        ("background", C.c_float), ("adu_per_nmgy", C.c_float),
        ("noise_additive", C.c_float), ("noise_multiplicative", C.c_float),
        ("normal_switch_rate", C.c_float),
    ]


class PriorParams(C.Structure):
    _fields_ = [
        ("count_kind", C.c_int32), ("flux_kind", C.c_int32),
        ("min_objects", C.c_int32), ("max_objects", C.c_int32),
        ("count_rate", C.c_float),
        ("loc_low", C.c_float * 2), ("loc_high", C.c_float * 2),
        ("flux_alpha", C.c_float), ("flux_lower", C.c_float), ("flux_upper", C.c_float),
        ("flux_logpdf_const", C.c_float),
        ("flux_mean", C.c_float), ("flux_stdev", C.c_float),
    ]


class MHParams(C.Structure):
    _fields_ = [
        ("num_iters", C.c_int32),
        ("locs_stdev", C.c_float), ("fluxes_stdev", C.c_float),
        ("fluxes_min", C.c_float), ("fluxes_max", C.c_float),
        ("locs_min", C.c_float * 2), ("locs_max", C.c_float * 2),
        ("refresh_loglik", C.c_int32),
        ("live_only", C.c_int32),
        ("acc_as_count", C.c_int32),
        ("live_tiles_hint", C.c_int32),
        ("tile_of_segment", C.c_void_p),
    ]


class LoopState(C.Structure):
    _fields_ = [("active_next", C.c_void_p), ("live_count", C.c_void_p), ("acc_count", C.c_void_p),
                ("acc_rate", C.c_void_p)]


class DrawTape(C.Structure):
    _fields_ = [("comp", C.c_void_p), ("u_loc", C.c_void_p), ("u_flux", C.c_void_p), ("u_acc", C.c_void_p)]


class MHTrace(C.Structure):
    _fields_ = [("log_alpha", C.c_void_p), ("target_prop", C.c_void_p), ("accept", C.c_void_p),
                ("chain_locs", C.c_void_p), ("chain_fluxes", C.c_void_p)]


class ResampledSource(C.Structure):
    _fields_ = [("index", C.c_void_p), ("counts", C.c_void_p), ("locs", C.c_void_p), ("fluxes", C.c_void_p),
                ("counts_out", C.c_void_p), ("copy_mask", C.c_void_p), ("rates", C.c_void_p),
                ("rates_out", C.c_void_p)]


_P = C.c_void_p
_I = C.c_int

PROTOTYPES = {
    "smcdet_version": (C.c_int, []),
    "smcdet_last_error_string": (C.c_char_p, []),
    "smcdet_loglik": (C.c_int, [C.POINTER(ModelParams), _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "smcdet_loglik_segments": (C.c_int, [C.POINTER(ModelParams), _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "smcdet_psf": (C.c_int, [C.POINTER(ModelParams), _P, _P, _I, _I, _I, _I, _I, _P]),
    "smcdet_psf_radial": (C.c_int, [C.POINTER(ModelParams), _I, _P, _P, C.c_longlong, _P]),
    "smcdet_render": (C.c_int, [C.POINTER(ModelParams), _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "smcdet_prior_logprob": (C.c_int, [C.POINTER(PriorParams), _P, _P, _P, _P, _I, _I, _I, _P]),
    "smcdet_prior_sample": (C.c_int, [C.POINTER(PriorParams), _P, _P, C.c_uint64, _P, _P, _P, _P, _I, _I, _I, _P]),
    "smcdet_temper_update": (C.c_int, [_P, _P, _P, C.c_float, _I, _P, _P, _P, _P, _P, _P, C.POINTER(LoopState), _I, _I, _P]),
    "smcdet_resample": (C.c_int, [_I, _P, _P, C.c_uint64, _P, _P, _P, _P, _I, _I, _P]),
    "smcdet_gather": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "smcdet_mh_mutate": (C.c_int, [C.POINTER(ModelParams), C.POINTER(PriorParams), C.POINTER(MHParams),
                                   _P, _P, _P, _P, _P, _P, _P, C.POINTER(DrawTape), C.POINTER(MHTrace),
                                   C.c_uint64, C.c_uint64, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "smcdet_mh_mutate_resampled": (C.c_int, [C.POINTER(ModelParams), C.POINTER(PriorParams), C.POINTER(MHParams),
                                             _P, C.POINTER(ResampledSource), _P, _P, _P, _P, _P, C.POINTER(DrawTape),
                                             C.POINTER(MHTrace), C.c_uint64, C.c_uint64, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "smcdet_mala_mutate": (C.c_int, [C.POINTER(ModelParams), C.POINTER(PriorParams), C.POINTER(MHParams),
                                     _P, _P, _P, _P, _P, _P, _P, C.POINTER(DrawTape), C.POINTER(MHTrace),
                                     C.c_uint64, C.c_uint64, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "smcdet_prune": (C.c_int, [_P, _P, C.c_float, C.c_float, C.c_float, _P, _P, _P, _I, _I, _I, _P]),
    "smcdet_agg_join": (C.c_int, [_P, _P, _I, C.c_float, _P, _P, _P, _I, _I, _I, _I, _P]),
    "smcdet_agg_unjoin": (C.c_int, [_P, _P, _I, C.c_float, _P, _P, _P, _I, _I, _I, _P]),
    "smcdet_agg_mutate": (C.c_int, [C.POINTER(ModelParams), C.POINTER(PriorParams), C.POINTER(MHParams), _I]
                          + [_P] * 10 + [C.POINTER(DrawTape), C.POINTER(MHTrace), C.c_uint64, C.c_uint64, _P, _P]
                          + [_I] * 5 + [_P]),
    "smcdet_match_catalogs": (C.c_int, [_P] * 8 + [C.c_float, C.c_float] + [_P] * 5 + [_I] * 6 + [_P]),
    "smcdet_debug_force_tpp": (C.c_int, [_I]),
}


def bind(cdll):
    """Attach restype/argtypes for every symbol the header declares; raises if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll
