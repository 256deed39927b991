"""ctypes binding of libdftbp_b200.so (include/dftbp_b200.h).

There is no fallback: if the CUDA library has not been built, importing the product API fails
loudly with the build instruction.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libdftbp_b200.so")

c_int, c_double, c_void_p, c_char = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_char
P = ctypes.POINTER

# name -> (restype, argtypes); every symbol the header declares
SIGNATURES = {
    "dftbp_b200_create": (c_int, [P(c_void_p), c_int, c_int, c_int]),
    "dftbp_b200_destroy": (c_int, [c_void_p]),
    "dftbp_b200_last_error": (ctypes.c_char_p, [c_void_p]),
    "dftbp_b200_version": (ctypes.c_char_p, []),
    "dftbp_b200_set_stream": (c_int, [c_void_p, c_void_p]),
    "dftbp_b200_get_stream": (c_void_p, [c_void_p]),
    "dftbp_b200_synchronize": (c_int, [c_void_p]),
    "dftbp_b200_set_orbitals": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dftbp_b200_set_sk_table": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_void_p, c_int]),
    "dftbp_b200_get_sk_cutoff": (c_int, [c_void_p, P(c_double)]),
    "dftbp_b200_set_geometry": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "dftbp_b200_cell_translations": (c_int, [c_void_p, c_void_p, c_double, c_int, P(c_int), c_void_p, c_void_p]),
    "dftbp_b200_fold_coords_to_cell": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "dftbp_b200_update_neighbour_list": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                                 c_double, c_double, c_int, P(c_int), P(c_int), P(c_int)]),
    "dftbp_b200_get_neighbour_list": (c_int, [c_void_p] + [c_void_p] * 10),
    "dftbp_b200_adopt_neighbour_list": (c_int, [c_void_p]),
    "dftbp_b200_nr_of_neighbours": (c_int, [c_void_p, c_double, c_void_p]),
    "dftbp_b200_update_coords": (c_int, [c_void_p, c_void_p]),
    "dftbp_b200_build_h0": (c_int, [c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_build_s": (c_int, [c_void_p, c_void_p]),
    "dftbp_b200_build_h0_s": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_unpack_hs_real": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "dftbp_b200_unpack_hs_cplx": (c_int, [c_void_p, c_void_p, P(c_double), c_void_p, c_int]),
    "dftbp_b200_pack_hs_real": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "dftbp_b200_pack_hs_cplx": (c_int, [c_void_p, c_void_p, c_void_p, c_int, P(c_double), c_double]),
    "dftbp_b200_density_matrix_real": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dftbp_b200_density_matrix_cplx": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dftbp_b200_density_matrix_real_cols": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                                    c_int, c_int]),
    "dftbp_b200_density_matrix_cplx_cols": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                                    c_int, c_int]),
    "dftbp_b200_pack_hs_real_atoms": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
    "dftbp_b200_pack_hs_cplx_atoms": (c_int, [c_void_p, c_void_p, c_void_p, c_int, P(c_double), c_double, c_int, c_int]),
    "dftbp_b200_mulliken_orbital": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_mulliken_block": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_mulliken_atom": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_onsite_population": (c_int, [c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_add_shift_atom": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "dftbp_b200_add_shift_block": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "dftbp_b200_derivative_shift": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double]),
    "dftbp_b200_set_pair_partition": (c_int, [c_void_p, c_int, c_int]),
    "dftbp_b200_stress": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double, c_double]),
    "dftbp_b200_hegv_real": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_char]),
    "dftbp_b200_hegv_cplx": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_char]),
    "dftbp_b200_timer_enable": (c_int, [c_void_p, c_int]),
    "dftbp_b200_timer_read": (c_int, [c_void_p, c_int, P(c_double), P(ctypes.c_longlong)]),
    "dftbp_b200_kernel_launches": (ctypes.c_longlong, [c_void_p]),
    "dftbp_b200_num_pairs": (ctypes.c_longlong, [c_void_p]),
    "dftbp_b200_efilling": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_int, c_double, c_double, c_void_p, c_int]),
    "dftbp_b200_invr_cluster": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int]),
    "dftbp_b200_invr_periodic": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_double, c_double,
                                         c_double, c_void_p, c_int]),
    "dftbp_b200_coulomb_shift": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "dftbp_b200_short_gamma_update": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_short_gamma_get": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_short_gamma_shifts": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "dftbp_b200_mixer_init": (c_int, [c_void_p, c_int, c_double, c_double, c_double, c_double, c_double]),
    "dftbp_b200_mixer_reset": (c_int, [c_void_p, c_int]),
    "dftbp_b200_mixer_mix": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
}

_lib = None


def load():
    """Load the CUDA library; raises if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "libdftbp_b200.so is not built (%s). Run `python -m dftbplus_b200.build` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
