"""CPU: the C-ABI library builds, loads, and exports every symbol that include/spn_b200.h declares, with the argument
counts the ctypes binding expects.  No compute call is made (there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "spn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\bint(?:64_t)?\s+(spn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def test_header_declares_the_expected_entry_points():
    names = set(declared_functions())
    for required in ("spn_leaf_fwd", "spn_leaf_bwd_params", "spn_leaf_bwd_x", "spn_sum_fwd", "spn_sum_bwd_w",
                     "spn_sum_bwd_x", "spn_xprod_fwd", "spn_xprod_bwd", "spn_sample_sum", "spn_sample_leaf",
                     "spn_version"):
        assert required in names


def test_library_exports_every_declared_symbol_with_matching_arity():
    import ratspn_b200 as rb
    lib = rb._abi.lib()
    decl = declared_functions()
    assert set(decl) == set(rb._abi.SIGNATURES), (set(decl) ^ set(rb._abi.SIGNATURES))
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} is declared in include/spn_b200.h but not exported"
        assert len(rb._abi.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, binding {len(rb._abi.SIGNATURES[name])}"
    assert "sm_100a" in rb._abi.version()


def test_library_is_built_for_sm100a_only():
    import subprocess
    import ratspn_b200 as rb
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", rb._abi.LIB_PATH], capture_output=True,
                         text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_product_fails_loudly_on_cpu_tensors():
    import torch
    import ratspn_b200 as rb
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 8, 2, 2, 2, 2, 1
    cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
    m = rb.RatSpn(cfg)
    with pytest.raises(rb._abi.SpnAbiError):
        m(torch.randn(2, 1, 8, 1, 1))
    with pytest.raises(rb._abi.SpnAbiError):
        m.sample(mode="index", n=2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "conditional-ratspn-for-rl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn


def test_tensor_core_leaf_shape_gate_is_a_host_function():
    """spn_leaf_tc_supported / spn_leaf_tc_image_bytes run on the host (no kernel): the gate of ops.LeafFn's tensor-core branch."""
    import ratspn_b200 as rb
    lib = rb._abi.lib()
    assert lib.spn_leaf_tc_supported(40, 40, 25, 32) == 1          # BASELINE config 2
    assert lib.spn_leaf_tc_supported(64, 64, 16, 64) == 1          # config 5 as named
    assert lib.spn_leaf_tc_supported(33, 2, 25, 4) == 0            # channels not a multiple of 8
    assert lib.spn_leaf_tc_supported(40, 40, 26, 32) == 0          # more than 25 features per scope (5 operand rows each, M = 128)
    assert lib.spn_leaf_tc_supported(72, 2, 8, 4) == 0             # more than 64 channels
    assert lib.spn_leaf_tc_image_bytes(40, 40, 25, 32) == 32 * 40 * 25 * 48 * 32 + 256
