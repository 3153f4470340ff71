"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/iisan_b200.h declares, struct layouts agree, and the Python mirror has the reference's parameter ABI."""
import ctypes
import os
import re

import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "iisan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iisan_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from iisan_b200 import _lib
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 15
    raw = ctypes.CDLL(_lib.lib_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/iisan_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    assert lib.iisan_abi_version() == _lib.ABI_VERSION
    assert lib.iisan_status_string(3).decode() == "workspace too small"


def test_struct_layouts_match():
    from iisan_b200 import _lib
    lib = _lib.load()
    for which, st in enumerate((_lib.SanDesc, _lib.SanParams, _lib.UeDesc, _lib.UeParams, _lib.CeDesc)):
        assert lib.iisan_sizeof(which) == ctypes.sizeof(st)


def test_descriptor_validation_without_gpu():
    """Pure host-side argument checks (no kernel is launched)."""
    from iisan_b200 import _lib
    lib = _lib.load()
    d = _lib.SanDesc()
    assert lib.iisan_san_workspace_bytes(ctypes.byref(d)) == 0            # empty descriptor rejected
    ce = _lib.CeDesc()
    ce.row_users, ce.col_users, ce.seq_len, ce.emb, ce.user_offset = 4, 2, 10, 64, 0     # rows outside the pool
    assert lib.iisan_inbatch_ce_workspace_bytes(ctypes.byref(ce)) == 0
    ce.col_users = 8
    assert lib.iisan_inbatch_ce_workspace_bytes(ctypes.byref(ce)) > 0
    assert lib.iisan_linear_forward(0, 1, 1, None, 0, None, None, None, 0, 0, None) == 1  # IISAN_EINVAL


@pytest.mark.parametrize("asym", [False, True])
def test_parameter_abi_matches_reference(asym):
    """names, order and shapes == the reference's named_parameters() (tests/golden were generated after asserting
    that oracle.synthetic.param_shapes equals the reference's own list)."""
    from oracle.synthetic import PathConfig, make_args, param_shapes
    if asym:
        from iisan_b200 import model_asym as pkg
        cfg = PathConfig(asym=True, d_text=96, d_img=64, layers_text=9, layers_img=5, bert_list="1,3,5,7", vit_list="1,3",
                         r_cv=16, r_bert=24, embedding_dim=32, item_num=500)
    else:
        from iisan_b200 import model as pkg
        cfg = PathConfig()
    args = make_args(cfg)

    class Img(nn.Module):
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(cfg.d_img, cfg.embedding_dim)

    m = pkg.ModelMM(args, cfg.item_num, True, Img(), nn.Identity(), [1.0] * (cfg.item_num + 1))
    m.mm_encoder = pkg.IISANAdaptedMModel(m.mm_encoder, args)
    got = [(n, tuple(p.shape)) for n, p in m.named_parameters()]
    assert got == list(param_shapes(cfg).items())
    if not asym:
        assert len(got) == 146 and sum(p.numel() for p in m.parameters()) == 4113877


def test_product_refuses_cpu_tensors():
    """No CPU fallback: the hot path raises instead of computing on the host."""
    from iisan_b200 import _lib
    from iisan_b200.model.modules import FusedLinear
    lin = FusedLinear(8, 4)
    with pytest.raises(_lib.IisanLibraryError):
        lin(torch.zeros(2, 8))


def test_unsupported_configs_rejected_at_construction():
    from oracle.synthetic import PathConfig, make_args
    from iisan_b200.plan import make_plan
    a = make_args(PathConfig()); a.fusion_method = "sum"
    with pytest.raises(NotImplementedError):
        make_plan(a, False)
    a = make_args(PathConfig()); a.modality = "intra"
    with pytest.raises(NotImplementedError):
        make_plan(a, False)


def test_gelu_activation_plan_descriptor_and_routing():
    """args.adapter_activation: nn.GELU() iff the string is exactly "GELU" (CC/model/modules.py:104-107).  The plan carries it
    into the descriptor; a GELU configuration is never routed to the ReLU-only fused chain kernels; the workspace grows by
    the pre-activation stash the GELU backward needs.  Host-side checks only (no kernel is launched)."""
    import torch
    from oracle.synthetic import PathConfig, make_args
    from iisan_b200 import _lib
    from iisan_b200.model.modules import AdapterBlock
    from iisan_b200.plan import SanBinder, make_plan
    lib = _lib.load()
    out = {}
    for act in ("RELU", "GELU", "gelu"):
        cfg = PathConfig(adapter_activation=act)
        args = make_args(cfg)
        plan = make_plan(args, False)
        assert plan.activation == (1 if act == "GELU" else 0)
        assert isinstance(AdapterBlock(args, 768, 64).activate, nn.GELU if act == "GELU" else nn.ReLU)
        binder = SanBinder(plan, [])
        img = torch.empty(2, 11, 13, 768, dtype=torch.float32); txt = torch.empty(2, 11, 13, 768, dtype=torch.float32)
        d = binder.desc(img, txt, _lib.COMPUTE_BF16, False)
        assert d.activation == plan.activation
        out[act] = (lib.iisan_san_fused_eligible(ctypes.byref(d)), lib.iisan_san_workspace_bytes(ctypes.byref(d)))
        d32 = binder.desc(img, txt, _lib.COMPUTE_FP32, False)
        out[act + "/fp32"] = lib.iisan_san_workspace_bytes(ctypes.byref(d32))
        d.activation = 7
        assert lib.iisan_san_workspace_bytes(ctypes.byref(d)) == 0           # unknown activation code rejected
        d.activation = plan.activation
    assert out["RELU"][0] == 1 and out["gelu"][0] == 1 and out["GELU"][0] == 0
    n, r, stages, towers = 22, 64, 7, 3
    assert out["GELU"][1] > out["RELU"][1] - 1 and out["GELU/fp32"] - out["RELU/fp32"] >= n * r * 4 * stages * towers


def test_grad_arena_hands_out_zeroed_aligned_slices():
    """ops.GradArena (one memset per step instead of a fill per backward function): slices are 256-byte aligned, disjoint, zero
    after reset(), and reset() clears exactly what the previous step used."""
    import torch
    from iisan_b200.ops import GradArena, grad_zeros
    a = GradArena(1024, "cpu")
    x = a.take(10); y = a.take(100)
    assert x.numel() == 10 and y.numel() == 100 and (y.data_ptr() - x.data_ptr()) == 64 * 4 and a.used().numel() == 192
    x.fill_(1.0); y.fill_(2.0)
    a.reset()
    assert a.off == 0 and float(a.buf.abs().sum()) == 0.0
    z = a.take(5)
    assert z.data_ptr() == x.data_ptr() and float(z.abs().sum()) == 0.0
    with pytest.raises(_lib_error()):
        a.take(2048)
    prev, GradArena.active = GradArena.active, a
    try:
        g = grad_zeros(7, "cpu")
        assert g.untyped_storage().data_ptr() == a.buf.untyped_storage().data_ptr()
    finally:
        GradArena.active = prev
    assert grad_zeros(7, "cpu").untyped_storage().data_ptr() != a.buf.untyped_storage().data_ptr()


def _lib_error():
    from iisan_b200._lib import IisanLibraryError
    return IisanLibraryError


def test_stage_plan_matches_oracle_plan():
    from oracle.iisan_oracle import stage_plan
    from oracle.synthetic import PathConfig, make_args
    from iisan_b200.plan import make_plan
    cfgs = [PathConfig(), PathConfig(remove_first="TRUE", bert_list="0,2,4,6,8,10"),
            PathConfig(asym=True, d_text=96, d_img=64, layers_text=9, layers_img=5, bert_list="1,3,5,7", vit_list="1,3"),
            PathConfig(asym=True, d_text=64, d_img=128, layers_text=5, layers_img=7, bert_list="1,3", vit_list="0,2,3,5")]
    for cfg in cfgs:
        plan = make_plan(make_args(cfg), cfg.asym)
        exp = [tuple(-1 if v is None else v for v in st) for st in stage_plan(cfg)]
        assert plan.stages == exp


def _versa_cfgs():
    from golden_util import VERSA_CASES, load_case
    from oracle.synthetic import PathConfig
    return {name: PathConfig(**load_case(name)[1]["cfg"]) for name in VERSA_CASES}


def test_versa_real_shapes_parameter_abi_plan_and_descriptors():
    """BASELINE configs[3]/[4] at their real widths: parameter names / shapes / order == the reference's (the fixtures were
    generated after asserting param_shapes == the reference's named_parameters()), the stage plan == the oracle's, and the C
    ABI accepts the descriptors of both arithmetic modes (host-side validation + workspace size only: no kernel launch)."""
    from oracle.iisan_oracle import stage_plan
    from oracle.synthetic import make_args, param_shapes
    from iisan_b200 import _lib, model_asym as pkg
    from iisan_b200.plan import SanBinder
    lib = _lib.load()
    expected_params = {"versa_llama70b_evaclip": 337_800_000}
    for name, cfg in _versa_cfgs().items():
        args = make_args(cfg)

        class Img(nn.Module):
            def __init__(self):
                super().__init__()
                self.classifier = nn.Linear(cfg.d_img, cfg.embedding_dim)

        with torch.device("meta"):                         # shapes only: the LLaMA/EVA case has 338 M parameters
            m = pkg.ModelMM(args, cfg.item_num, True, Img(), nn.Identity(), [1.0] * 4)
            m.mm_encoder = pkg.IISANAdaptedMModel(m.mm_encoder, args)
        got = [(n, tuple(p.shape)) for n, p in m.named_parameters()]
        assert got == list(param_shapes(cfg).items()), name
        total = sum(p.numel() for p in m.parameters())
        if name in expected_params:
            assert abs(total - expected_params[name]) < 0.01 * expected_params[name], total     # SURVEY 8a row a9 [probe]
        plan = m.mm_encoder.plan
        assert plan.stages == [tuple(-1 if v is None else v for v in st) for st in stage_plan(cfg)], name
        binder = SanBinder(plan, [n for n, _ in m.mm_encoder.named_parameters()])
        n_items = 22
        for dtype in (torch.float32, torch.bfloat16, torch.float16):
            image = torch.empty(n_items, cfg.layers_img, cfg.d_img, dtype=dtype, device="meta")
            text = torch.empty(n_items, cfg.layers_text, cfg.d_text, dtype=dtype, device="meta")
            for compute in (_lib.COMPUTE_FP32, _lib.COMPUTE_BF16):
                d = binder.desc(image, text, compute)
                assert d.n_stages == len(plan.stages) and d.d_mm == min(cfg.d_text, cfg.d_img)
                nbytes = lib.iisan_san_workspace_bytes(ctypes.byref(d))
                assert 0 < nbytes < (4 << 30), (name, dtype, compute, nbytes)
        # a selected layer outside the cached states is refused on the host
        with pytest.raises(_lib.IisanLibraryError):
            binder.desc(torch.empty(n_items, 3, cfg.d_img, device="meta"), torch.empty(n_items, 3, cfg.d_text, device="meta"),
                        _lib.COMPUTE_FP32)


@pytest.mark.parametrize("tree", ["Code_Cached", "Code_Cached_Asym"])
def test_state_dict_equals_reference_checkpoint_layout(tree, tmp_path):
    """SURVEY 8f-4 / 8b "Parameter ABI": ``state_dict()`` of the drop-in model has the keys, shapes, dtypes and order of the
    reference model's (tests/golden/state_dict_keys.json, frozen from the reference's own ModelMM + IISANAdaptedMModel by
    oracle/make_golden_statedict.py), so an ``epoch-N.pt`` (utils.py:104-110: model_state_dict / optimizer / rng_state /
    cuda_rng_state) written by either side loads into the other (run.py:234-243)."""
    import json
    from oracle.synthetic import PathConfig, make_args
    z = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))[tree]
    cfg = PathConfig(**z["cfg"])
    if cfg.asym:
        from iisan_b200 import model_asym as pkg
    else:
        from iisan_b200 import model as pkg
    args = make_args(cfg)

    class Img(nn.Module):
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(cfg.d_img, cfg.embedding_dim)

    def build():
        m = pkg.ModelMM(args, 50, True, Img(), nn.Identity(), [1.0] * 51)
        m.mm_encoder = pkg.IISANAdaptedMModel(m.mm_encoder, args)
        return m

    m = build()
    got = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert got == z["entries"]
    # the reference's checkpoint dict, written here and loaded the way run.py:234-243 does (strict key matching)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    path = tmp_path / "epoch-3.pt"
    torch.save({"model_state_dict": m.state_dict(), "optimizer": opt.state_dict(), "rng_state": torch.get_rng_state(),
                "cuda_rng_state": torch.ByteTensor(8)}, path)
    ck = torch.load(path, map_location=torch.device("cpu"))
    m2 = build()
    missing, unexpected = m2.load_state_dict(ck["model_state_dict"])
    assert not missing and not unexpected
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.parametrize("tree", ["Code_Cached", "Code_Cached_Asym"])
def test_lr_groups_equal_reference_routing(tree):
    """the learning rate of every parameter as routed by the reference's own training script (the substring tests
    of Code_Cached/run.py:260-307, executed from /root/reference by oracle/make_golden_lr_groups.py over the reference model's
    parameter names and frozen in tests/golden/lr_groups.json) equals what iisan_b200.optim.param_groups assigns -- the groups
    FusedAdam / torch.optim.Adam are built from."""
    import argparse
    import json
    from iisan_b200.optim import param_groups
    z = json.load(open(os.path.join(ROOT, "tests", "golden", "lr_groups.json")))[tree]
    names = list(z["lr_of"].keys())

    class Param:
        requires_grad = True

    params = {n: Param() for n in names}

    class Named:
        def named_parameters(self):
            return list(params.items())

    groups = param_groups(Named(), argparse.Namespace(**z["lrs"]))
    assert [len(g["params"]) for g in groups] == z["group_sizes"]
    got = {}
    for g in groups:
        for p in g["params"]:
            name = next(n for n, q in params.items() if q is p)
            assert name not in got
            got[name] = g["lr"]
    assert got == z["lr_of"]
