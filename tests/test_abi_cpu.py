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
    assert lib.molclr_abi_version() == 1
    assert lib.molclr_plan_workspace_bytes(10, 20, 3) >= 4 * 23
    assert lib.molclr_gemm_colstat_tiles(129) == 8 and lib.molclr_gemm_colstat_tile_rows() == 32


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


def test_gemm_args_struct_matches_header_layout():
    from molclr_b200._lib import GemmArgs
    # 8-byte aligned fields in header order; guards against silent drift between header and ctypes mirror
    assert GemmArgs.A.offset == 0 and GemmArgs.lda.offset == 8 and GemmArgs.a_mn.offset == 16
    assert GemmArgs.B.offset == 24 and GemmArgs.A_lo.offset == 48 and GemmArgs.M.offset == 64 and GemmArgs.out.offset == 88
    assert GemmArgs.out_lo.offset == 128 and ctypes.sizeof(GemmArgs) == 240 and GemmArgs.relu_bits.offset == 208 and GemmArgs.compensate.offset == 232


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
