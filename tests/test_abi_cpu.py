"""CPU tests: the C-ABI library builds, loads and exports every symbol include/molclr_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from molclr_b200.build import build
    return build()


def test_library_exports_every_declared_symbol(lib_path):
    hdr = open(os.path.join(ROOT, "include", "molclr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(molclr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from molclr_b200 import _lib
    assert set(_lib.SIGNATURES) == declared


def test_library_loads_and_reports_version(lib_path):
    from molclr_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "molclr_b200.h")).read()
    assert lib.molclr_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define MOLCLR_ABI_VERSION (\d+)", hdr).group(1))
    assert lib.molclr_plan_workspace_bytes(10, 20, 3) >= 4 * 23
    assert lib.molclr_gemm_colstat_tiles(129) == 8 and lib.molclr_gemm_colstat_tile_rows() == 32


def test_programmatic_launch_switch_is_a_plain_host_toggle(lib_path):
    """molclr_set_pdl: default on; returns the previous value; a negative argument only queries."""
    from molclr_b200 import _lib
    lib = _lib.load()
    first = lib.molclr_set_pdl(-1)
    try:
        assert first in (0, 1)
        assert lib.molclr_set_pdl(0) == first and lib.molclr_set_pdl(-1) == 0
        assert lib.molclr_set_pdl(1) == 0 and lib.molclr_set_pdl(-1) == 1
        assert lib.molclr_set_pdl(7) == 1 and lib.molclr_set_pdl(-1) == 1          # any non-zero value means "on"
    finally:
        lib.molclr_set_pdl(first)


def test_ntxent_workspace_covers_both_backward_variants(lib_path):
    """Host-only sizing: one buffer serves the forward partials, the fp16 operand copies, the striped backward (W stripe +
    per-stripe partial gradients + cols^T) and the fused backward (32 split slots + column factors), for any shape."""
    from molclr_b200 import _lib
    lib = _lib.load()
    for R, Rc, C in ((8, 8, 16), (12, 12, 12), (8192, 8192, 256), (8192, 65536, 256), (600, 600, 320)):
        n = lib.molclr_ntxent_workspace_bytes(R, Rc, C)
        ld16 = (C + 7) // 8 * 8
        fused = 32 * R * C * 4 + (R + Rc) * ld16 * 2 + (Rc + 63) // 64 * 64 * 4
        striped = R * 8192 + ((Rc + 2047) // 2048) * R * C * 4 + (R + Rc) * ld16 * 2 + C * ((Rc + 7) // 8 * 8) * 2
        assert n >= fused and n >= striped, (R, Rc, C, n, fused, striped)
        assert lib.molclr_ntxent_workspace_bytes(R, 2 * Rc, C) > n


def test_struct_mirrors_match_the_header_as_compiled_by_gcc(tmp_path):
    """The ctypes mirrors of every struct of the header (GEMM arguments, weight descriptors, model / layer / plan views) have exactly the layout a C compiler gives the header's structs
    (guards against silent drift between header and binding)."""
    import subprocess
    from molclr_b200._lib import GemmArgs, WeightDesc, GinLayer, GinModel, PlanView
    mirrors = {"molclr_gemm_args": GemmArgs, "molclr_weight_desc": WeightDesc, "molclr_gin_layer": GinLayer, "molclr_gin_model": GinModel,
               "molclr_plan_view": PlanView}
    fields = {st: [f[0] for f in cls._fields_] for st, cls in mirrors.items()}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "molclr_b200.h"', 'int main(void) {']
    for st, fs in fields.items():
        src.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        src += [f'  printf("{st}.{f} %zu\\n", offsetof({st}, {f}));' for f in fs]
    src += ['  return 0;', '}']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for st, cls in mirrors.items():
        assert int(got[st]) == ctypes.sizeof(cls), st
        for f in fields[st]:
            assert int(got[f"{st}.{f}"]) == getattr(cls, f).offset, (st, f)


def test_product_has_no_cpu_fallback():
    import molclr_b200
    m = molclr_b200.GINet(2, 32, 16)
    from molclr_b200.synth import make_pair_batch
    bi, _ = make_pair_batch(2, seed=0)
    with pytest.raises(RuntimeError):
        m(bi)
    # nothing under the package imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "molclr_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_drop_in_constructor_signatures():
    import inspect
    import molclr_b200
    sig = inspect.signature(molclr_b200.GINet.__init__)
    assert list(sig.parameters)[1:] == ["num_layer", "emb_dim", "feat_dim", "drop_ratio", "pool"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [5, 300, 256, 0, "mean"]
    sig = inspect.signature(molclr_b200.NTXentLoss.__init__)
    assert list(sig.parameters)[1:] == ["device", "batch_size", "temperature", "use_cosine_similarity"]


def test_gcn_drop_in_layout_and_errors():
    """GCN state_dict keys/shapes/dtypes == the reference's shipped checkpoint (tests/golden/gcn_ckpt_manifest.json)."""
    import json
    import molclr_b200
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "gcn_ckpt_manifest.json")))["entries"]
    sd = molclr_b200.GCN(5, 300, 512, 0, "mean").state_dict()
    assert set(sd.keys()) == set(man.keys())
    for k, v in sd.items():
        assert list(v.shape) == man[k]["shape"] and str(v.dtype).replace("torch.", "") == man[k]["dtype"], k
    with pytest.raises(ValueError):
        molclr_b200.GCN(1, 32, 16)
    with pytest.raises(ValueError):
        molclr_b200.GCN(3, 32, 16, 0, "bogus")
    import inspect
    sig = inspect.signature(molclr_b200.GCN.__init__)
    assert list(sig.parameters)[1:] == ["num_layer", "emb_dim", "feat_dim", "drop_ratio", "pool"]
