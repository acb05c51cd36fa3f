"""mebt_b200.pipelines (bidirect_sample / extrapolate of the reference's sample_vqgan_transformer_videos.py:22-157)
against what the reference's own functions produced when driven by the same deterministic stand-in model
(tests/golden/make_golden.py::gen_pipelines): identical code maps, scores, pixel post-processing and - call by call -
identical arguments to `model.sample`.  CPU only: the pipelines are host orchestration."""
import json

import numpy as np
import torch

from conftest import load_golden
from helpers import FakeSampler


def _check(tag, z, model, log):
    assert torch.equal(log["code_maps"], torch.from_numpy(z[f"{tag}_code_maps"]))
    np.testing.assert_allclose(log["samples"].numpy(), z[f"{tag}_samples"], rtol=0, atol=0)
    if f"{tag}_score" in z.files:
        np.testing.assert_allclose(log["score"].numpy(), z[f"{tag}_score"], rtol=1e-6)
    assert len(model.calls) == int(z[f"{tag}_n_calls"])
    for i, c in enumerate(model.calls):
        assert torch.equal(c["x"], torch.from_numpy(z[f"{tag}_call{i}_x"])), (tag, i)
        assert torch.equal(c["ctx"], torch.from_numpy(z[f"{tag}_call{i}_ctx"])), (tag, i)
        assert torch.equal(c["tgt"], torch.from_numpy(z[f"{tag}_call{i}_tgt"])), (tag, i)
        meta = json.loads(str(z[f"{tag}_call{i}_meta"]))
        mine = {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items() if k not in ("x", "ctx", "tgt")}
        assert mine == meta, (tag, i, mine, meta)


def test_bidirect_sample_matches_reference_script():
    from mebt_b200.pipelines import bidirect_sample
    z, _ = load_golden("pipelines")
    m = FakeSampler((4, 4, 4))
    log = bidirect_sample(m, 2, total_length=16, step_size=16, context_size=12, temperature=0.9, top_k=5, vid_n_steps=6,
                          vid_c_temp=2.0, ctemp_schedule="cosine", strategy="maskgit", bootstrap=3,
                          decode=m.first_stage_model.decode)
    _check("bi_one", z, m, log)
    m = FakeSampler((4, 4, 4))
    log = bidirect_sample(m, 3, total_length=16, step_size=16, context_size=8, vid_n_steps=4, strategy="random",
                          decode=m.first_stage_model.decode)
    _check("bi_one_nb", z, m, log)


def test_bidirect_sample_sliding_windows():
    """More than one window: the context frames of window j are the tail of window j-1, targets are the rest; the score
    gather then fails exactly as in the reference (first-window probabilities vs the full code map)."""
    from mebt_b200 import pipelines
    m = FakeSampler((4, 4, 4))
    try:
        pipelines.bidirect_sample(m, 2, total_length=40, step_size=16, context_size=8, vid_n_steps=4)
        raised = False
    except RuntimeError:
        raised = True
    assert raised
    assert len(m.calls) == 4                                   # first window + ceil((10 - 4) / 2) = 3 slides
    hw = 16
    for c in m.calls[1:]:
        assert torch.equal(c["ctx"], torch.arange(2 * hw).repeat(2, 1))
        assert torch.equal(c["tgt"], torch.arange(2 * hw).repeat(2, 1) + 2 * hw)
        assert not c["edit"] and not c["debug"]
        assert int(c["x"][:, 2 * hw:].abs().sum()) == 0        # everything but the context is forgotten


def test_extrapolate_matches_reference_script():
    from mebt_b200.pipelines import extrapolate
    z, _ = load_golden("pipelines")
    vq = torch.from_numpy(z["ex_input"])
    m = FakeSampler((4, 4, 4))
    log = extrapolate(m, vq, total_length=40, step_size=16, context_size=8, temperature=0.8, top_p=0.9, vid_n_steps=5,
                      vid_c_temp=3.0, decode=m.first_stage_model.decode)
    _check("ex", z, m, log)
    assert all(c["edit"] for c in m.calls)
    m = FakeSampler((4, 4, 4))
    log = extrapolate(m, vq, total_length=30, step_size=16, context_size=12, vid_n_steps=7, decode=m.first_stage_model.decode)
    _check("ex_odd", z, m, log)
    # the default decoder is the model's own first stage (the script's gpt.first_stage_model.decode); a token-only model
    # (vtokens: True, no first stage) returns the token-level log
    default = extrapolate(FakeSampler((4, 4, 4)), vq, total_length=30, step_size=16, context_size=12, vid_n_steps=7)
    assert torch.equal(default["code_maps"], log["code_maps"]) and torch.equal(default["samples"], log["samples"])
    tokens_only = FakeSampler((4, 4, 4))
    tokens_only.first_stage_model = None
    no_pixels = extrapolate(tokens_only, vq, total_length=30, step_size=16, context_size=12, vid_n_steps=7)
    assert torch.equal(no_pixels["code_maps"], log["code_maps"]) and no_pixels["samples"] == []
