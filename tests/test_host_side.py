"""Host-side pieces that need no GPU: the NUMA placement helper and the exports of the C++ level driver's library."""
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpulist_parsing():
    from phare_b200 import numa
    assert numa.parse_cpulist("0-3,8,10-11") == [0, 1, 2, 3, 8, 10, 11]
    assert numa.parse_cpulist("") == [] and numa.parse_cpulist(None) == []
    assert numa.parse_cpulist("5") == [5]


def test_preferred_node_syscall_is_reachable():
    """set_mempolicy(MPOL_PREFERRED, node 0) through the raw syscall: node 0 exists on every Linux box"""
    from phare_b200 import numa
    assert numa.set_preferred_node(0) in (True, False)  # never raises; True wherever the syscall is permitted


def test_host_library_exports_the_level_driver():
    lib_path = os.path.join(ROOT, "phare_b200", "lib", "libphare_b200_host.so")
    if not os.path.exists(lib_path):
        pytest.skip("libphare_b200_host.so not built (make host)")
    from phare_b200 import abi
    abi.load()  # its symbols resolve against libphare_b200.so
    lib = ctypes.CDLL(lib_path)
    for name in ("phh_create", "phh_create_distributed", "phh_arena_export", "phh_arena_open", "phh_release_peers",
                 "phh_patch_id", "phh_destroy", "phh_ctx", "phh_npatch", "phh_layout", "phh_field", "phh_add_population",
                 "phh_particles", "phh_initialize", "phh_advance", "phh_last_error"):
        assert hasattr(lib, name), name
