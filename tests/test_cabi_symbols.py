"""CPU: libomc.so loads without a GPU and exports exactly the symbols include/omc.h declares; the ctypes prototype
table covers every one of them; the product package never imports the oracle."""

import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "omc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(omc_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_exported_and_bound():
    from openmcmc_b200 import _cabi

    lib = _cabi.load()
    declared = _declared()
    assert declared, "no declarations parsed from include/omc.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/omc.h but not exported by libomc.so"
    assert declared == set(_cabi.PROTOTYPES), (declared ^ set(_cabi.PROTOTYPES))
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (omc_[a-z0-9_]+)", out))
    assert declared <= exported
    assert lib.omc_abi_version() == 1


def test_struct_sizes_match_header():
    """ctypes mirrors of the argument structs have the sizes the C compiler gives them."""
    import ctypes
    import tempfile

    from openmcmc_b200 import _cabi

    names = {"omc_vec_t": _cabi.Vec, "omc_rng_t": _cabi.Rng, "omc_nn_dense_t": _cabi.NNDense,
             "omc_quadform_t": _cabi.Quadform, "omc_ng_draw_t": _cabi.NGDraw, "omc_term_t": _cabi.Term,
             "omc_mh_model_t": _cabi.MHModel, "omc_random_walk_t": _cabi.RandomWalkArgs, "omc_mmala_t": _cabi.MMalaArgs,
             "omc_linear_predictor_t": _cabi.LinearPredictor, "omc_logp_gamma_t": _cabi.LogpGamma,
             "omc_logp_poisson_t": _cabi.LogpPoisson, "omc_logp_normal_ss_t": _cabi.LogpNormalSS}
    names.update(getattr(_cabi, "EXTRA_STRUCTS", {}))
    src = '#include <stdio.h>\n#include "omc.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert ctypes.sizeof(names[n]) == int(sz), (n, ctypes.sizeof(names[n]), sz)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "openmcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in text, os.path.join(dirpath, f)


def test_no_gpu_means_loud_failure():
    import pytest
    import torch

    from openmcmc_b200 import _cabi
    from openmcmc_b200 import kernels as K

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_cabi.OmcError):
        K.init_device()
