"""CPU-side checks of the drop-in boundary: libb200zk.so builds, loads and exports every symbol that
include/b200zk.h declares; with no GPU the library refuses to initialise (no CPU fallback)."""
import ctypes as C

import noir_backend_using_gnark_b200 as zk


def test_exports_every_declared_symbol(lib_built):
    lib = zk.load()
    names = zk.header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    # the binding covers the whole header
    assert set(lib._signatures) == set(names)


def test_strerror_and_null_handling(lib_built):
    lib = zk.load()
    assert lib.b200zk_strerror(0) == b"ok"
    assert b"CPU fallback" in lib.b200zk_strerror(-1)
    assert lib.b200zk_init(0, None) == -3
    assert lib.b200zk_sync(None) == -3
    assert lib.b200zk_ntt_dev(None, None, 4, 0, 0, 0) == -3
    assert lib.b200zk_launch_count(None) == 0
    lib.b200zk_destroy(None)
    lib.b200zk_bases_free(None, None)
    # every later addition to the ABI refuses null handles the same way
    assert lib.b200zk_plonk_prove_hex(None, None, None, 0, None, None) == -3
    assert lib.b200zk_plonk_set_solution_map(None, None, None, 0) == -3
    assert lib.b200zk_plonk_set_commit_lanes(None, 3) == -3
    assert lib.b200zk_plonk_unsatisfied_row(None) == -1
    assert lib.b200zk_msm_windows(None, None, 0) == -3
    assert lib.b200zk_host_alloc(None, 16, None) == -3
    assert lib.b200zk_msm_set_reduce_chunk(None, 3) == -3
    assert lib.b200zk_msm_set_small_path(None, 1) == -3
    assert lib.b200zk_msm_set_host_chunks(None, 2) == -3
    assert lib.b200zk_msm_set_pair_rounds(None, 2) == -3
    assert lib.b200zk_msm_set_scatter_passes(None, 2) == -3
    assert b"constraint" in lib.b200zk_strerror(-6)


def test_no_cpu_fallback(lib_built):
    lib = zk.load()
    if lib.b200zk_device_count() > 0:
        return  # on a GPU box the gpu-marked tests cover initialisation
    h = C.c_void_p()
    assert lib.b200zk_init(0, C.byref(h)) == -1 and not h
    try:
        zk.Context(0)
    except zk.B200zkError as e:
        assert e.code == -1
    else:  # pragma: no cover
        raise AssertionError("Context() must fail without a CUDA device")


def test_product_does_not_import_oracle():
    import pathlib

    pkg = pathlib.Path(zk.__file__).parent
    for p in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*.cu*")) + list((pkg / "csrc").glob("*.h")) + list((pkg / "csrc" / "ffi").glob("*")):
        text = p.read_text()
        assert "oracle" not in text.lower(), p
