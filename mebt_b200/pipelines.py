"""Token-level generation pipelines of the reference's sampling script (sample_vqgan_transformer_videos.py):

* `bidirect_sample` (:22-93)  - optional `bootstrap` pass, a maskgit pass over the first window, then sliding windows
  whose first `context_size` latent frames are the tail of what has been generated so far;
* `extrapolate`     (:95-157) - the same sliding window started from given tokens, with `edit=True` (the mask schedule
  counts only the editable tokens, mebt/transformer.py:373-376).

They are host orchestration over `Net2NetTransformer.sample` (whose forwards, sampling and re-masking run on the CUDA
kernels) and return the reference's `log` dict at token level: `code_maps`, `class_label` and, for `bidirect_sample`,
`score`.  Pixels (`log["samples"]`) come from the VQGAN decoder exactly as the script computes them
(`gpt.first_stage_model.decode`, :82,146): by default the model's own first stage (`mebt_b200.vqgan.VQGAN`, loaded by
`init_first_stage_from_ckpt` when `vtokens: False`); `decode=<callable>` overrides it, and a token-only model
(`vtokens: True`, no first stage) returns the token-level log.

Reference quirks kept: the temporal ratio 0.25 is hard-coded (:29,104); `skips=False` is passed positionally as in the
script; the windows after the first one in `bidirect_sample` are sampled WITHOUT `edit` (:62) while `extrapolate` uses
`edit=True` (:137); `score` gathers the first window's probabilities with the full code map (:88-90), so it only exists
when the whole video fits one window - otherwise `torch.gather` raises, as it does in the reference.
"""
from __future__ import annotations

import numpy as np
import torch


def _decode(model, code_map, decode, total_length, log):
    if decode is None:
        first = getattr(model, "first_stage_model", None)
        decode = first.decode if first is not None else None
    if decode is None:
        return
    img_x = decode(code_map)
    log["samples"] = (torch.clamp(img_x, -0.5, 0.5) + 0.5)[:, :, :total_length, :, :]


@torch.no_grad()
def bidirect_sample(model, batch_size, total_length, step_size, context_size, temperature=1.0, top_k=None, top_p=None,
                    frame_n_steps=8, vid_n_steps=8, frame_c_temp=4.5, vid_c_temp=4.5, no_phase=False,
                    ctemp_schedule="linear", strategy="maskgit", bootstrap=0, decode=None):
    T, H, W = model.mask_sampler.shape[-3:]
    ratio = 0.25
    step_size = int(step_size * ratio)
    context_size = int(context_size * ratio)
    shape = (batch_size, step_size, H, W)
    device = model.device
    log = dict(samples=[])
    log["class_label"] = torch.zeros(batch_size, 1, dtype=torch.long, device=device)
    code_map = []

    x = torch.zeros(shape, dtype=torch.long, device=device)
    context_indices = target_indices = None
    bs_partial_probs = None
    # the script only ever gathers the dense [B, N, 16384] probability maps at the final codes: ask the sampler for that
    # gather directly when it can provide it (mebt_b200's Net2NetTransformer), the dense maps otherwise
    extra = dict(debug_probs="selected") if getattr(model, "selected_probs_supported", False) else {}
    if bootstrap > 0:
        x, context_indices, target_indices, _, _, bs_partial_probs = model.sample(
            x, None, 1., None, None, bootstrap, context_indices, target_indices, context_temperature=vid_c_temp,
            skips=False, ctemp_schedule=ctemp_schedule, strategy="bootstrap", debug=True, **extra)
    x, context_indices, _, _, _, final_partial_probs = model.sample(
        x, None, temperature, top_k, top_p, vid_n_steps, context_indices, target_indices,
        context_temperature=vid_c_temp, skips=False, ctemp_schedule=ctemp_schedule, strategy=strategy, debug=True, **extra)
    curr_t = step_size
    vq_x = x.reshape(shape)
    code_map.append(vq_x)

    hw = H * W
    while curr_t < total_length * ratio:
        new_x = torch.zeros(shape, dtype=torch.long, device=device)          # forget everything but the context frames
        new_x[:, :context_size] = vq_x[:, -context_size:]
        context_indices = torch.arange(hw * context_size, device=device).repeat(batch_size, 1)
        target_indices = torch.arange((step_size - context_size) * hw, device=device).repeat(batch_size, 1) + hw * context_size
        x = model.sample(new_x, None, temperature, top_k, top_p, vid_n_steps, context_indices, target_indices,
                         context_temperature=vid_c_temp, skips=False, ctemp_schedule=ctemp_schedule, strategy=strategy)[0]
        vq_x = x.reshape(shape)
        code_map.append(vq_x[:, context_size:])
        curr_t += step_size - context_size
    code_map = torch.cat(code_map, 1)
    if code_map.shape[1] == 1:
        code_map = code_map.expand(-1, 4, H, W)
    _decode(model, code_map, decode, total_length, log)
    log["code_maps"] = code_map
    if bs_partial_probs is not None:
        final_prob_map = torch.where(final_partial_probs < 0., bs_partial_probs, final_partial_probs)
    else:
        final_prob_map = final_partial_probs
    if final_prob_map.dim() == 2:          # already the probability of each position's final code
        selected = final_prob_map
    else:
        selected = torch.gather(final_prob_map, -1, code_map.reshape(batch_size, -1, 1)).squeeze(-1)
    log["score"] = selected.log().sum(-1)
    return log


@torch.no_grad()
def extrapolate(model, vq_input, total_length, step_size, context_size, temperature=1.0, top_k=None, top_p=None,
                frame_n_steps=8, vid_n_steps=8, frame_c_temp=4.5, vid_c_temp=4.5, no_phase=False,
                ctemp_schedule="linear", strategy="maskgit", bootstrap=0, decode=None):
    B, T, H, W = vq_input.shape
    ratio = 0.25
    step_size = int(step_size * ratio)
    context_size = int(context_size * ratio)
    assert T == step_size
    total_size = int(total_length * ratio)
    jump_size = step_size - context_size
    n_jumps = int(np.ceil((total_size - step_size) / jump_size))
    device = model.device
    log = dict(samples=[])
    log["class_label"] = torch.zeros(B, 1, dtype=torch.long, device=device)
    code_map = [vq_input.clone()]

    indices = torch.arange(H * W * step_size, device=device).repeat(B, 1).view(B, step_size, H, W)
    context_indices = indices[:, :context_size].reshape(B, -1)
    target_indices = indices[:, context_size:].reshape(B, -1)

    x = vq_input
    for _ in range(n_jumps):
        window = torch.zeros_like(x)                                          # forget everything but the context frames
        window[:, :context_size] = code_map[-1][:, -context_size:]
        x = model.sample(window.view(B, -1), None, temperature, top_k, top_p, vid_n_steps, context_indices,
                         target_indices, context_temperature=vid_c_temp, skips=False, edit=True)[0]
        x = x.view(B, step_size, H, W)
        code_map.append(x.clone()[:, context_size:])
    code_map = torch.cat(code_map, 1)
    _decode(model, code_map, decode, total_length, log)
    log["code_maps"] = code_map
    return log
