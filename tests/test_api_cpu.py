"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares; the drop-in classes
expose the reference's API surface and state_dict; host-side mask logic matches the reference fixtures; and
there is no silent CPU fallback."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import REPO, load_golden
from helpers import build_model, model_configs


def test_library_exports_every_declared_symbol():
    header = (REPO / "include" / "mebt_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(mebt_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 20
    lib = ctypes.CDLL(str(REPO / "mebt_b200" / "libmebt_b200.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    lib.mebt_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mebt_version()


def test_ctypes_signatures_cover_header():
    from mebt_b200 import _lib
    header = (REPO / "include" / "mebt_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    for name, args in re.findall(r"\bint\s+(mebt_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S):
        n_args = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert name in _lib._SIGNATURES, name
        assert len(_lib._SIGNATURES[name]) == n_args, (name, len(_lib._SIGNATURES[name]), n_args)


def test_no_gpu_means_loud_failure():
    from mebt_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib.mebt_device_check() != 0 and _lib.last_error()
    with pytest.raises(_lib.MebtError):
        ops.layernorm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))
    z, cfg = load_golden("forward_micro")
    model = build_model(cfg, None, device="cpu")
    x = torch.zeros(1, 256, dtype=torch.long)
    idx = torch.arange(256).view(1, -1)
    with pytest.raises(_lib.MebtError):
        model.reconstruct_mask(x, idx[:, :10], idx[:, 10:])


def test_state_dict_matches_reference_contract():
    from oracle import mebt_oracle as O
    z, cfg = load_golden("forward_tiny")
    model = build_model(cfg, None, device="cpu")
    sd = model.state_dict()
    shapes = O.param_shapes(cfg)
    assert set(sd.keys()) == set(shapes.keys())
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == tuple(shp) and sd[k].dtype == torch.float32, k
    # registration order starts mask_emb, sos_emb, pos_emb, transformer..., tok_emb
    names = [n for n, _ in model.named_parameters()]
    assert names[:3] == ["mask_emb", "sos_emb", "pos_emb"] and names[-1] == "tok_emb.weight"
    # a reference-style state_dict loads strictly
    model.load_state_dict(O.make_weights(cfg, 1), strict=True)


def test_api_surface():
    import inspect

    import mebt
    import mebt.modules.gpt
    import mebt.transformer as T
    from mebt_b200.mask_sampler import MaskGen
    from mebt_b200.modules.codebook import Codebook
    from mebt_b200.modules.gpt import GPT, Block, CrossAttention
    N2N = T.Net2NetTransformer
    for name in ("forward", "reconstruct_mask", "sample", "entp_sample", "draft", "revise", "draft_and_revise",
                 "shared_step", "training_step", "validation_step", "configure_optimizers", "optimizer_step",
                 "encode_to_z", "encode_to_c", "get_input", "get_xc", "top_k_logits", "init_from_ckpt",
                 "add_model_specific_args"):
        assert callable(getattr(N2N, name)), name
    for name in ("sample_from_logits", "gumbel_sort", "top_k_logits", "top_p_probs", "uniform", "gaussian", "gaussian2",
                 "gaussian100000_2", "longest", "linear", "constant", "cosine"):
        assert callable(getattr(T, name)), name
    sig = inspect.signature(N2N.draft_and_revise)
    assert list(sig.parameters)[1:] == ["x", "c", "n_draft", "draft_t", "draft_k", "draft_p", "n_revise", "revise_t",
                                        "revise_k", "revise_p", "M", "skip_draft", "debug", "context_indices",
                                        "target_indices", "edit"]
    sig = inspect.signature(N2N.sample)
    assert list(sig.parameters)[1:9] == ["x", "c", "temperature", "top_k", "top_p", "n_steps", "context_indices",
                                         "target_indices"]
    assert sig.parameters["context_temperature"].default == 4.5 and sig.parameters["strategy"].default == "maskgit"
    assert list(inspect.signature(GPT.__init__).parameters)[1:] == [
        "vocab_size", "block_size", "n_layer", "n_head", "n_embd", "embd_pdrop", "resid_pdrop", "attn_pdrop",
        "n_unmasked", "vtokens_pos", "mode"]
    assert list(inspect.signature(MaskGen.__init__).parameters)[1:] == ["iid", "schedule", "max_token", "method", "shape",
                                                                        "t_range", "budget"]
    assert list(inspect.signature(Codebook.__init__).parameters)[1:] == ["n_codes", "embedding_dim", "no_random_restart",
                                                                         "restart_thres"]
    assert mebt.Net2NetTransformer is N2N and mebt.modules.gpt.Block is Block and CrossAttention is not None
    with pytest.raises(ValueError):
        MaskGen(schedule="nope")
    # t_prior / video length priors
    L = np.arange(32) + 1
    assert T.longest(L, 0)[-1] == 1.0 and T.longest(L, 0).sum() == 1.0 and T.uniform(L, 0).sum() == 32
    assert np.argmax(T.gaussian2(L, 30000 * 9)) == 9 and np.argmax(T.gaussian100000_2(L, 0)) == 0


def test_maskgen_host_logic_matches_reference_fixture():
    from mebt_b200.mask_sampler import MaskGen
    z, _ = load_golden("maskgen")
    N = 1024
    for sched in ("cosine", "linear", "quadratic", "sqrt", "square", "cube", "cosine_plus", "convex"):
        mg = MaskGen(schedule=sched, shape=(4, 16, 16), budget=1024).eval()
        g = torch.Generator().manual_seed(5)
        indices = torch.stack([torch.randperm(N, generator=g) for _ in range(2)])
        sizes = []
        for t in z["ts"]:
            c, tg, sl = mg.divide_indices(indices, torch.tensor(float(t)), None, None)
            sizes.append([c.shape[1], tg.shape[1], int(sl)])
        assert (np.array(sizes) == z[f"{sched}_sizes"]).all(), sched
        n_masked = []
        for steps in (8, 32, 128):
            for t_next in np.linspace(0, 1, steps + 1)[1:]:
                n_masked.append(float(torch.ceil(mg.schedule_fn(torch.full((2,), fill_value=t_next)) * N)[0]))
        assert (np.array(n_masked) == z[f"{sched}_n_masked"]).all(), sched
    # training-mode frame-window slicing consumes the numpy RNG like the reference
    mg = MaskGen(schedule="linear", shape=(4, 16, 16), budget=300).train()
    np.random.seed(3)
    c, tg, sl = mg.divide_indices(torch.from_numpy(z["train_indices"]), torch.tensor(0.4), np.arange(4) + 1,
                                  np.array([0.1, 0.2, 0.3, 0.4]))
    assert int(sl) == int(z["train_meta"][0])
    assert (c.numpy() == z["train_ctx"]).all() and (tg.numpy() == z["train_tgt"]).all()
    # gibbs partitions draw randperm from the CPU generator
    e, a = torch.empty(2, 0).long(), torch.arange(N).repeat(2, 1)
    torch.manual_seed(78)
    cs, ts = MaskGen.create_gibbs_draft_mask(e, a, 4, "cpu")
    assert (cs[3].numpy() == z["draft_ctx3"]).all() and (ts[3].numpy() == z["draft_tgt3"]).all()
    torch.manual_seed(78)
    cs, ts = MaskGen.create_gibbs_revise_mask(e, a, 4, "cpu")
    assert (cs.numpy() == z["revise_ctx"]).all() and (ts.numpy() == z["revise_tgt"]).all()
    with pytest.raises(AssertionError):
        MaskGen.create_gibbs_revise_mask(e, a, 3, "cpu")


def test_configure_optimizers_groups():
    z, cfg = load_golden("forward_micro")
    model = build_model(cfg, None, device="cpu")
    model.learning_rate, model.weight_decay = 1e-4, 0.01
    opt = model.configure_optimizers()
    g = opt.param_groups
    assert len(g) == 4 and g[0]["weight_decay"] == 0.01 and all(x["weight_decay"] == 0.0 for x in g[1:])
    n_layer = cfg["n_layer"]
    assert len(g[0]["params"]) == 6 * n_layer + 1            # q,k,v,proj,fc1,fc2 weights per block + head
    assert len(g[1]["params"]) == 3                          # mask_emb, sos_emb, tok_emb.weight
    assert len(g[2]["params"]) == 10 * n_layer + 2           # 6 biases + 4 LayerNorm params per block + ln_f
    assert len(g[3]["params"]) == 1 and g[3]["params"][0] is model.pos_emb
    assert opt.defaults["betas"] == (0.9, 0.95)
    total = sum(len(x["params"]) for x in g)
    assert total == len(list(model.parameters()))


def test_config_yaml_shape_parses():
    """The shipped STL config's structure (configs/stl/mebt_16f.yaml) drives construction unchanged."""
    from helpers import STL_16F
    cfg = dict(STL_16F, n_embd=128, n_head=2, sos_emb=32)     # same 24-entry mode list, small width for CPU
    params, vq, mask = model_configs(cfg)
    from mebt_b200.transformer import Net2NetTransformer
    m = Net2NetTransformer(params, vq, mask)
    modes = [b.mode for b in m.transformer.blocks]
    assert modes.count("latent_enc") == 7 and modes.count("latent_self") == 6
    assert modes.count("latent_dec") == 6 and modes.count("lt2l") == 5 and modes[-1] == "latent_dec"
    assert m.mask_sampler.schedule == "linear" and m.t_prior.__name__ == "longest"


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/mebt_b200.h must compile as C99 (no torch / C++ types in signatures)."""
    import shutil
    import subprocess
    from pathlib import Path
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    hdr = Path(__file__).resolve().parent.parent / "include" / "mebt_b200.h"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", str(hdr)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import re
    code = re.sub(r"/\*.*?\*/", "", hdr.read_text(), flags=re.S)          # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "#include <c" not in code.replace("#include <cuda", "")


def test_first_stage_vqgan_is_loaded_and_frozen_when_vtokens_is_false(tmp_path):
    """`vtokens: False` (transformer.py:180-192): the VQGAN of first_stage_config.params.ckpt_path becomes the frozen
    `first_stage_model` (reference key names in the state_dict, no gradients, train() is a no-op on it, vocabulary size from its
    codebook), and the sampling-script pipelines decode through it by default."""
    import torch
    from helpers import STL_16F, to_attr
    from mebt_b200 import vqgan as V
    from mebt_b200.transformer import Net2NetTransformer
    vcfg = dict(embedding_dim=64, n_codes=128, n_hiddens=32, downsample=(2, 4, 4), image_channels=3, norm_type="group",
                padding_type="replicate", sequence_length=8, sample_every_n_frames=1, resolution=32)
    vq = V.VQGAN(V._Args(vcfg))
    ckpt = tmp_path / "vqgan.ckpt"
    torch.save({"hyper_parameters": {"args": V._Args(vcfg)}, "state_dict": vq.state_dict()}, ckpt)
    cfg = dict(STL_16F, n_embd=64, n_head=1, sos_emb=16, n_layer=4, mode=["latent_enc", "latent_self", "latent_dec", "lt2l"],
               vocab_size=128, block_size=256, shape=[4, 8, 8])
    params, _, mask = model_configs(cfg)
    params.vtokens = False
    first = to_attr(dict(params=dict(ckpt_path=str(ckpt), ignore_keys=["loss"])))
    m = Net2NetTransformer(params, first, mask)
    assert isinstance(m.first_stage_model, V.VQGAN) and m.first_stage_vocab_size == 128
    assert m.first_stage_model.latent_shape == (4, 8, 8)
    assert all(not p.requires_grad for p in m.first_stage_model.parameters())
    assert m.first_stage_model.codebook._need_init is False
    m.train()
    assert m.transformer.training and not m.first_stage_model.training and not m.first_stage_model.encoder.training
    keys = m.state_dict().keys()
    assert "first_stage_model.encoder.conv_first.conv.weight" in keys and "first_stage_model.codebook.embeddings" in keys
    assert torch.equal(m.first_stage_model.state_dict()["decoder.conv_last.conv.weight"], vq.state_dict()["decoder.conv_last.conv.weight"])
    # the pipelines' default decoder is the model's own first stage
    from mebt_b200 import pipelines
    seen = {}
    m.first_stage_model.decode = lambda codes: (seen.setdefault("codes", codes), torch.zeros(codes.shape[0], 3, 8, 32, 32))[1]
    log = {}
    pipelines._decode(m, torch.zeros(1, 4, 8, 8, dtype=torch.long), None, 8, log)
    assert "codes" in seen and log["samples"].shape == (1, 3, 8, 32, 32)


def test_small_utility_helpers():
    """mebt/utils.py:55-124,174-178 and mebt/modules/gpt.py:19-42: generic helpers the reference's scripts import."""
    import torch
    from mebt_b200.modules.gpt import complement_idx
    from mebt_b200.utils import accuracy, adopt_weight, comp_getattr, correct, tensor_slice, view_range
    x = torch.arange(2 * 24 * 3).view(2, 24, 3)
    assert view_range(x, 1, 2, (2, 3, 4)).shape == (2, 2, 3, 4, 3) and view_range(x, -2, -1, (4, 6)).shape == (2, 4, 6, 3)
    assert view_range(x, 1, None, (8, 9)).shape == (2, 8, 9)
    assert torch.equal(tensor_slice(x, (0, 4, 1), (-1, 5, 2)), x[:, 4:9, 1:3])
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.15, 0.05], [0.2, 0.3, 0.5]])
    tgt = torch.tensor([1, 1, 0])
    c1, c2 = correct(out, tgt, topk=(1, 2))
    assert float(c1) == 1.0 and float(c2) == 2.0
    a1, a2 = accuracy(out, tgt, topk=(1, 2))
    assert abs(float(a1) - 100.0 / 3) < 1e-4 and abs(float(a2) - 200.0 / 3) < 1e-4
    assert adopt_weight(10, threshold=50, value=0.0) == 0.0 and adopt_weight(50, threshold=50) == 1
    assert comp_getattr(type("A", (), {"k": 3})(), "k") == 3 and comp_getattr(object(), "k", 7) == 7
    g = torch.Generator().manual_seed(0)
    idx = torch.stack([torch.stack([torch.randperm(10, generator=g)[:4] for _ in range(3)]) for _ in range(2)])     # [2, 3, 4]
    comp = complement_idx(idx, 10)
    assert comp.shape == (2, 3, 6)
    for b in range(2):
        for r in range(3):
            assert comp[b, r].tolist() == sorted(set(range(10)) - set(idx[b, r].tolist()))
    assert torch.equal(complement_idx(torch.empty(2, 0, dtype=torch.long), 5), torch.arange(5).repeat(2, 1))


def test_load_transformer_from_a_lightning_style_checkpoint(tmp_path):
    """mebt/download.py:56-61: the model is re-created from the checkpoint's `hyper_parameters` (the constructor arguments
    `save_hyperparameters()` records) and its `state_dict` is loaded strictly."""
    import torch
    from helpers import STL_16F
    from mebt.download import load_transformer
    from mebt_b200._lib import MebtError
    from mebt_b200.transformer import Net2NetTransformer
    cfg = dict(STL_16F, n_embd=64, n_head=1, sos_emb=16, n_layer=4, mode=["latent_enc", "latent_self", "lt2l", "latent_dec"],
               vocab_size=128, block_size=256, shape=[4, 8, 8])
    params, vq, mask = model_configs(cfg, schedule="cosine")
    torch.manual_seed(3)
    m = Net2NetTransformer(params, vq, mask)
    ckpt = tmp_path / "gpt.ckpt"
    torch.save({"state_dict": m.state_dict(), "hyper_parameters": dict(transformer_config=params, first_stage_config=vq, mask_config=mask,
                                                                      ckpt_path="/nonexistent/previous/run.ckpt", first_stage_key="video",
                                                                      cond_stage_key="label", pkeep=1.0, sos_token=0)}, ckpt)
    got = load_transformer(str(ckpt), "ignored_vqgan.ckpt")
    assert not got.training and got.mask_sampler.schedule == "cosine"
    a, b = m.state_dict(), got.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    try:
        load_transformer({"state_dict": {}})
        raised = False
    except MebtError:
        raised = True
    assert raised


def test_out_of_scope_modules_resolve_to_the_reference_checkout(tmp_path, monkeypatch):
    """`from mebt import VideoData`, `from mebt.data import preprocess`, `from mebt.utils import save_video_grid` (the imports
    of the reference's scripts) resolve to `$MEBT_REF/mebt/{data,utils}.py`; without MEBT_REF they fail with a message naming it."""
    import importlib
    import sys
    for k in [k for k in sys.modules if k in ("mebt.data",) or k.startswith("mebt._reference_")]:
        del sys.modules[k]
    monkeypatch.delenv("MEBT_REF", raising=False)
    import mebt
    try:
        mebt.VideoData
        raised = ""
    except ImportError as exc:
        raised = str(exc)
    assert "MEBT_REF" in raised
    try:
        from mebt.utils import save_video_grid  # noqa: F401
        ok = True
    except ImportError as exc:
        ok = "MEBT_REF" in str(exc) and False
    assert not ok
    fake = tmp_path / "MeBT" / "mebt"
    fake.mkdir(parents=True)
    (fake / "data.py").write_text("class VideoData:\n    tag = 'reference data module'\n\ndef preprocess(video, resolution):\n    return ('pre', resolution)\n")
    (fake / "utils.py").write_text("def save_video_grid(video, fname, nrow=None):\n    return ('saved', fname)\n")
    monkeypatch.setenv("MEBT_REF", str(tmp_path / "MeBT"))
    for k in [k for k in sys.modules if k in ("mebt.data",) or k.startswith("mebt._reference_")]:
        del sys.modules[k]
    from mebt import VideoData
    assert VideoData.tag == "reference data module"
    data = importlib.import_module("mebt.data")
    assert data.preprocess(None, 128) == ("pre", 128)
    from mebt.utils import save_video_grid, shift_dim
    assert save_video_grid(None, "x.mp4") == ("saved", "x.mp4") and callable(shift_dim)
    for k in [k for k in sys.modules if k in ("mebt.data",) or k.startswith("mebt._reference_")]:
        del sys.modules[k]
