"""ctypes binding of libb200pic.so (include/b200pic.h).

There is no CPU fallback: if the CUDA library is missing, or no CUDA device is
usable when compute is requested, this raises.
"""
import ctypes as C
import os

from ._abi import AntennaMode, B2PConfig, EdgeBC, ParticleState, ReflectorWall

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2P_LIB: a kernel-variant build of the same library (tools/microbench.py experiments); never a fallback
SO_PATH = os.environ.get("B2P_LIB") or os.path.join(_HERE, "libb200pic.so")

# every symbol include/b200pic.h declares (checked by tests/test_abi.py)
SYMBOLS = """
b2p_last_error b2p_version b2p_init b2p_sync b2p_set_option b2p_host_register b2p_host_unregister b2p_gpu_mem_kB
b2p_tile_create b2p_tile_destroy b2p_tile_bounds
b2p_tile_set_fields b2p_tile_get_fields b2p_tile_push_half_b b2p_tile_push_e b2p_tile_add_current
b2p_tile_filter_current b2p_tile_clear_current b2p_tile_field_energy
b2p_tile_inject b2p_tile_set_particles b2p_tile_container_size b2p_tile_get_particles
b2p_tile_push_particles b2p_tile_deposit_current b2p_tile_sort_particles b2p_tile_pack_outgoing_particles
b2p_tile_sort_keys b2p_tile_get_outgoing b2p_tile_kinetic_energy
b2p_tile_register_edge_bc b2p_tile_apply_edge_bcs b2p_tile_apply_edge_bc
b2p_tile_register_reflector_wall b2p_tile_reflect_particles b2p_tile_advance_reflector_walls b2p_tile_reflector_walls
b2p_tile_register_antenna b2p_tile_deposit_antenna_current
b2p_grid_apply_edge_bcs b2p_grid_reflect_particles b2p_grid_advance_reflector_walls b2p_grid_write_fields_snapshot b2p_grid_alive_counts
b2p_grid_create b2p_grid_destroy b2p_grid_add_tile b2p_grid_local_communication
b2p_grid_push_half_b b2p_grid_push_e b2p_grid_add_current b2p_grid_filter_current
b2p_grid_push_particles b2p_grid_pack_outgoing_particles b2p_grid_sort_particles b2p_grid_deposit_current
b2p_grid_step_pic b2p_grid_step_emf b2p_grid_energies b2p_grid_inject_thermal b2p_grid_inject_drifting_stripe b2p_grid_set_uniform_B
b2p_nccl_unique_id b2p_grid_comm_init b2p_grid_external_communication b2p_plan_describe
b2p_timer_start b2p_timer_stop b2p_launch_count b2p_copy_bytes b2p_host_wait_ms
b2p_profile_enable b2p_profile_num_classes b2p_profile_class_name b2p_profile_report
b2p_selfcheck_const_division
""".split()


class B2PError(RuntimeError):
    """std::runtime_error of the reference."""


class B2PLogicError(Exception):
    """std::logic_error of the reference (pybind11 maps it to a plain Exception subclass)."""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, ci, u64, dp = C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_double)
    L.b2p_last_error.restype = C.c_char_p
    L.b2p_version.restype = C.c_char_p
    L.b2p_gpu_mem_kB.restype = C.c_int64
    L.b2p_launch_count.restype = C.c_uint64
    L.b2p_init.argtypes = [ci]
    L.b2p_set_option.argtypes = [C.c_char_p, ci]
    L.b2p_host_register.argtypes = [vp, C.c_size_t]
    L.b2p_host_unregister.argtypes = [vp]
    L.b2p_tile_create.argtypes = [C.POINTER(B2PConfig), C.POINTER(C.c_int32 * 3), C.POINTER(vp)]
    L.b2p_tile_destroy.argtypes = [vp]
    L.b2p_tile_destroy.restype = None
    L.b2p_tile_bounds.argtypes = [vp, C.POINTER(C.c_double * 3), C.POINTER(C.c_double * 3)]
    L.b2p_tile_set_fields.argtypes = [vp, vp, vp, vp, ci]
    L.b2p_tile_get_fields.argtypes = [vp, vp, vp, vp, ci]
    for n in ("push_half_b", "push_e", "add_current", "filter_current", "clear_current", "push_particles",
              "deposit_current", "sort_particles", "pack_outgoing_particles"):
        getattr(L, "b2p_tile_" + n).argtypes = [vp]
    L.b2p_tile_field_energy.argtypes = [vp, dp, dp]
    L.b2p_tile_inject.argtypes = [vp, ci, u64] + [vp] * 6
    L.b2p_tile_set_particles.argtypes = [vp, ci, u64] + [vp] * 7
    L.b2p_tile_container_size.argtypes = [vp, ci, C.POINTER(u64)]
    L.b2p_tile_get_particles.argtypes = [vp, ci, ci] + [vp] * 7 + [C.POINTER(u64)]
    L.b2p_tile_sort_keys.argtypes = [vp, ci, vp]
    L.b2p_tile_get_outgoing.argtypes = [vp, vp, u64, vp, C.POINTER(u64)]
    L.b2p_tile_kinetic_energy.argtypes = [vp, ci, dp, C.POINTER(u64)]
    L.b2p_tile_register_edge_bc.argtypes = [vp, C.POINTER(EdgeBC)]
    L.b2p_tile_apply_edge_bcs.argtypes = [vp, ci]
    L.b2p_tile_apply_edge_bc.argtypes = [vp, C.POINTER(EdgeBC), ci]
    L.b2p_tile_register_reflector_wall.argtypes = [vp, C.POINTER(ReflectorWall)]
    L.b2p_tile_reflect_particles.argtypes = [vp]
    L.b2p_tile_advance_reflector_walls.argtypes = [vp]
    L.b2p_tile_reflector_walls.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.b2p_tile_register_antenna.argtypes = [vp, C.POINTER(AntennaMode)]
    L.b2p_tile_deposit_antenna_current.argtypes = [vp]
    L.b2p_grid_apply_edge_bcs.argtypes = [vp, ci]
    L.b2p_grid_reflect_particles.argtypes = [vp]
    L.b2p_grid_advance_reflector_walls.argtypes = [vp]
    L.b2p_grid_alive_counts.argtypes = [vp, vp]
    L.b2p_grid_write_fields_snapshot.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
    L.b2p_grid_create.argtypes = [C.POINTER(B2PConfig), C.POINTER(vp)]
    L.b2p_grid_destroy.argtypes = [vp]
    L.b2p_grid_destroy.restype = None
    L.b2p_grid_add_tile.argtypes = [vp, vp]
    L.b2p_grid_local_communication.argtypes = [vp, ci]
    for n in ("push_half_b", "push_e", "add_current", "filter_current", "push_particles",
              "pack_outgoing_particles", "sort_particles", "deposit_current", "step_emf"):
        getattr(L, "b2p_grid_" + n).argtypes = [vp]
    L.b2p_grid_step_pic.argtypes = [vp, C.c_int64]
    L.b2p_grid_energies.argtypes = [vp, dp, dp, vp, vp]
    L.b2p_grid_inject_thermal.argtypes = [vp, ci, C.c_double, u64]
    L.b2p_grid_inject_drifting_stripe.argtypes = [vp, ci, ci, C.c_double, C.c_double, ci, C.c_double, C.c_double, u64]
    L.b2p_grid_set_uniform_B.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    L.b2p_nccl_unique_id.argtypes = [vp]
    L.b2p_grid_comm_init.argtypes = [vp, ci, ci, vp, vp]
    L.b2p_grid_external_communication.argtypes = [vp, ci]
    L.b2p_plan_describe.argtypes = [C.POINTER(B2PConfig), vp, ci, vp, C.c_int64]
    L.b2p_plan_describe.restype = C.c_int64
    L.b2p_timer_stop.argtypes = [C.POINTER(C.c_float)]
    L.b2p_copy_bytes.argtypes = [C.POINTER(u64), C.POINTER(u64)]
    L.b2p_copy_bytes.restype = None
    L.b2p_host_wait_ms.argtypes = [C.POINTER(C.c_double)]
    L.b2p_host_wait_ms.restype = None
    L.b2p_profile_enable.argtypes = [ci]
    L.b2p_profile_class_name.argtypes = [ci]
    L.b2p_profile_class_name.restype = C.c_char_p
    L.b2p_profile_report.argtypes = [vp, vp, vp]
    L.b2p_selfcheck_const_division.argtypes = [vp, u64, C.c_float, vp, vp]
    _lib = L
    return L


def check(rc):
    if rc == 0:
        return
    msg = lib().b2p_last_error().decode(errors="replace")
    if rc == 2:
        raise B2PLogicError(msg)
    raise B2PError(msg)
