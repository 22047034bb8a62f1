"""CPU checks of the drop-in boundary: the library builds, loads, and exports every symbol that
include/socialways_b200.h declares (no compute calls here: there is no GPU on the CPU tier)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "socialways_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(sw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from socialways_b200 import build, _lib
    build.build()
    assert os.path.exists(_lib.LIB_PATH)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes prototypes out of sync with the header"
    lib = _lib.lib()
    assert lib.sw_abi_version() == 2
    assert lib.sw_decode_pack_floats() == 160 * 160 + 160 + 160 * 80 + 80 + 160 + 2
    assert lib.sw_pool_pack_floats() == 128 + 64 * 32 + 64
    assert b"argument" in lib.sw_error_string(-1)


def test_null_pointers_are_rejected_without_touching_the_gpu():
    from socialways_b200 import _lib
    lib = _lib.lib()
    assert lib.sw_decode_fwd(*([None] * 13), 1, 1, 1, 148, None) == -1
    assert lib.sw_pool_fwd(None, None, None, None, None, None, None, None, 1, 1, None) == -1
    assert lib.sw_bestofk_metrics(None, None, 1.0, 1, 1, 1, None, None) == -1
    assert lib.sw_lstm_seq_fwd(None, None, 2, 1, 8, *([None] * 8), 148, None) == -1


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "socialways_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", ""), fn


def test_no_cpu_fallback():
    import pytest
    import torch
    import socialways_b200 as sw
    g = sw.Generator(use_social=True)
    with pytest.raises(sw.SocialWaysCudaError):
        g.predict_k(torch.zeros(4, 8, 2), torch.zeros(1, 4, 32), 12, [[0, 4]])


def test_new_rows_have_no_cpu_path_either():
    """statistics / fused optimiser: CPU inputs are rejected, nothing falls back to numpy or torch.optim."""
    import numpy as np
    import pytest
    import torch
    import socialways_b200 as sw
    if torch.cuda.is_available():
        pytest.skip("CPU-tier check")
    with pytest.raises(sw.SocialWaysCudaError):
        sw.fused_optim.FlatAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    with pytest.raises((sw.SocialWaysCudaError, RuntimeError, AssertionError)):
        sw.statistics.compute_wasserstein(np.zeros((3, 2, 4, 2), np.float32), np.zeros((3, 2, 4, 2), np.float32))


def test_null_pointers_are_rejected_by_the_new_entry_points():
    from socialways_b200 import _lib
    lib = _lib.lib()
    assert lib.sw_traj_nn1_counts(None, None, 4, 1, 1, 1, 4, 2, None, None) == -1
    assert lib.sw_traj_emd_cost(None, None, 4, 1, 1, 4, 2, None, None) == -1
    assert lib.sw_lsap_solve(None, 4, 1, None, None, None) == -1
    assert lib.sw_adam_flat(None, None, None, None, None, 4, 1e-3, 0.9, 0.999, 1e-8, 148, None) == -1
    assert lib.sw_allreduce_adam(None, 0, 2, 4, 32, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, None) == -1
    assert lib.sw_pool_fwd_tcx(None, None, None, None, None, None, None, None, None, 0, 0, 0, 1, 1, None) == -1
    assert lib.sw_pool_tcx_max_scene() == 512 and lib.sw_lsap_smem_bytes(20) == 20 * 42


def test_training_step_entry_points_reject_bad_arguments():
    """Round-2 entry points (native training iteration): argument checks only, nothing touches a GPU."""
    import ctypes
    from socialways_b200 import _lib
    from socialways_b200.native_step import ContractJob
    lib = _lib.lib()
    assert lib.sw_gen_pack(None, *([None] * 7), None) == -1
    assert lib.sw_gen_pack_bwd(None, None, None, None, None, 1, None) == -1
    assert lib.sw_disc_pack(None, 48, None, None, None, None) == -1
    assert lib.sw_disc_step(*([None] * 1), 48, 0, None, None, None, None, 8, None, 32, None, 1.0, 0.5, *([None] * 7), 4, 148, None) == -1
    assert lib.sw_rows_linear(None, 64, None, None, None, None, None, 65, 4, 64, 65, None) == -1
    assert lib.sw_train_stats(None, None, 4, 12, 1.0, None, 0, None, 0, 1.0, 0.5, None, None, None, 148, None) == -1
    assert lib.sw_set_peer_wait_timeout_ms(0) == -1 and lib.sw_set_peer_wait_timeout_ms(60000) == 0
    sizes = [ctypes.c_int() for _ in range(7)]
    assert lib.sw_gen_pack_sizes(*[ctypes.byref(s) for s in sizes]) == 0
    assert [s.value for s in sizes] == [69 * 256, 256 * 68, 38802, 23200, 2240, 65 * 65, 65 * 64]
    xr, gr = ctypes.c_int(), ctypes.c_int()
    assert lib.sw_disc_step_image_rows(48, ctypes.byref(xr), ctypes.byref(gr)) == 0 and (xr.value, gr.value) == (304, 196)
    # contraction planning is host arithmetic: a job over many images is split into chunks with a partial-slab workspace
    from socialways_b200.native_step import ContractSeg
    assert ctypes.sizeof(ContractJob) == 176 and ctypes.sizeof(ContractSeg) == 32
    segs = (ContractSeg * 3)(ContractSeg(1 << 22, None, 0, 69, 256, 1))
    job = ContractJob(1 << 20, 1 << 21, 68 * 32, 256 * 32, 0, 68, 0, 256, 5000, 5000 * 32, 0, 0, 1, 0, 1, 0, segs)
    ws, nc = ctypes.c_longlong(), ctypes.c_int()
    assert lib.sw_contract_plan((ContractJob * 1)(job), 1, 148, 0, ctypes.byref(ws), ctypes.byref(nc)) == 0
    assert nc.value == 2 * 4 and ws.value > 0 and ws.value % 4096 == 0           # FFMA: 64 x 64 tiles of the 69 x 256 result
    assert lib.sw_contract_plan((ContractJob * 1)(job), 1, 148, 1, ctypes.byref(ws), ctypes.byref(nc)) == 0
    assert nc.value == 2 and ws.value > 0 and ws.value % (128 * 128) == 0        # tcgen05: two 128 x 128 slabs
    bad = ContractJob(None, None, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, (ContractSeg * 3)())
    assert lib.sw_contract_plan((ContractJob * 1)(bad), 1, 148, 1, ctypes.byref(ws), ctypes.byref(nc)) == -1
    assert lib.sw_contract_tc(None, 0, None, 0, None, 0, 148, None) == -1
