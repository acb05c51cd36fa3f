"""Shared test helpers: config objects and model construction for the drop-in package."""
import numpy as np
import torch


class AttrDict(dict):
    """Stands in for OmegaConf nodes: attribute access, hasattr, `in`, .get."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    return AttrDict({k: to_attr(v) for k, v in d.items()}) if isinstance(d, dict) else d


def model_configs(cfg: dict, schedule="linear"):
    params = to_attr(dict(
        unconditional=True, vocab_size=cfg["vocab_size"], first_stage_vocab_size=cfg["vocab_size"],
        block_size=cfg["block_size"], n_layer=cfg["n_layer"], n_head=cfg["n_head"], n_embd=cfg["n_embd"], n_unmasked=0,
        embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0, sample_every_n_latent_frames=0, first_stage_key="video",
        cond_stage_key="label", vtokens=True, vtokens_pos=False, vis_epoch=100, sos_emb=cfg["sos_emb"],
        avg_loss=bool(cfg.get("avg_loss", 1.0)), mode=list(cfg["mode"]), class_cond_dim=None))
    mask = to_attr(dict(target="mebt.mask_sampler.MaskGen",
                        params=dict(iid=False, schedule=schedule, max_token=cfg["block_size"], method="mlm",
                                    shape=cfg["shape"], t_range=[0.0, 1.0], budget=cfg["block_size"])))
    vq = to_attr(dict(params=dict(ckpt_path="unused", ignore_keys=["loss"])))
    return params, vq, mask


def build_model(cfg: dict, weights: dict | None = None, schedule="linear", device="cuda"):
    from mebt_b200.transformer import Net2NetTransformer
    params, vq, mask = model_configs(cfg, schedule)
    model = Net2NetTransformer(params, vq, mask)
    if weights is not None:
        missing, unexpected = model.load_state_dict(weights, strict=True)
        assert not missing and not unexpected
    return model.to(device).eval()


STL_16F = dict(n_embd=1024, n_head=16, sos_emb=256, block_size=1024, shape=[4, 16, 16], n_layer=24, vocab_size=16384,
               avg_loss=1.0,
               mode=["latent_enc", "latent_self"] * 6 + ["latent_enc"] + ["latent_dec", "lt2l"] * 5 + ["latent_dec"])
STL_128F = dict(STL_16F, block_size=8192, shape=[32, 16, 16])


def synth_tokens(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    N = int(np.prod(cfg["shape"]))
    x = torch.randint(0, cfg["vocab_size"], (B, *cfg["shape"]), generator=g)
    indices = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    return x, indices


def reference_dropout_masks(z, p):
    """Keep factors recorded from the reference's nn.Dropout calls (tests/golden/make_golden.py::gen_grads_dropout),
    keyed the way oracle.mebt_oracle.gpt_forward(drop=...) expects them."""
    site = {"attn.attn_drop": "attn", "attn.resid_drop": "proj", "mlp.3": "mlp"}
    drop, stem = {}, iter(("lat", "ctx", "tgt", "mask_emb"))
    for i, name in enumerate(str(n) for n in z["mask_names"]):
        shape = tuple(int(v) for v in z[f"mask_shape:{i}"])
        keep = np.unpackbits(z[f"mask_bits:{i}"])[:int(np.prod(shape))].reshape(shape)
        factor = torch.from_numpy(keep.astype(np.float32)) / (1.0 - p)
        if name == "transformer.drop":                 # GPT.forward drops sos_emb, contexts, targets, mask_emb in this order
            drop[("stem", next(stem))] = factor
        else:
            layer, sub = name[len("transformer.blocks."):].split(".", 1)
            drop[(int(layer), site[sub])] = factor
    return drop


class FakeSampler:
    """Stands in for Net2NetTransformer in the pipeline tests (bidirect_sample / extrapolate are host orchestration over
    `model.sample`): deterministic tokens, every call recorded.  Used both by tests/golden/make_golden.py (driving the
    REFERENCE's functions) and by tests/test_pipelines_cpu.py (driving mebt_b200.pipelines)."""

    V = 32

    class _MS:
        def __init__(self, shape):
            self.shape = shape

    class _FS:
        @staticmethod
        def decode(code_map):
            B, T, H, W = code_map.shape
            return (code_map.float().view(B, 1, T, H, W).repeat(1, 3, 4, 1, 1) / 16.0) - 1.0

    def __init__(self, shape):
        self.mask_sampler = self._MS(shape)
        self.first_stage_model = self._FS()
        self.device = torch.device("cpu")
        self.calls = []

    def sample(self, x, c, temperature=1.0, top_k=None, top_p=None, n_steps=8, context_indices=None, target_indices=None,
               strategy="maskgit", context_temperature=4.5, phase_history=None, refine_steps=1, forget_pivot=False,
               skips=(False, False, False), debug=False, ctemp_schedule="linear", edit=False):
        B = x.shape[0]
        x = x.reshape(B, -1).clone()
        N = x.shape[1]
        k = len(self.calls)
        if target_indices is None:
            g = torch.Generator().manual_seed(100 + k)
            perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
            context_indices, target_indices = perm[:, :0], perm
        self.calls.append(dict(temperature=temperature, top_k=top_k, top_p=top_p, n_steps=n_steps, strategy=strategy,
                               context_temperature=context_temperature, skips=skips, debug=debug,
                               ctemp_schedule=ctemp_schedule, edit=edit, x=x.clone(),
                               ctx=context_indices.clone(), tgt=target_indices.clone()))
        vals = (target_indices * 7 + 3 * k + 1) % self.V
        if strategy == "bootstrap":                       # reveals n_steps tokens, the rest stays masked
            n_new = min(n_steps, target_indices.shape[1])
            x.scatter_(1, target_indices[:, :n_new], vals[:, :n_new])
            ctx_out = torch.cat([context_indices, target_indices[:, :n_new]], 1)
            tgt_out = target_indices[:, n_new:]
            filled = target_indices[:, :n_new]
        else:
            x.scatter_(1, target_indices, vals)
            ctx_out = torch.cat([context_indices, target_indices], 1)
            tgt_out = target_indices[:, :0]
            filled = target_indices
        if not debug:
            return x, ctx_out, tgt_out
        probs = -torch.ones(B, N, self.V)
        g = torch.Generator().manual_seed(200 + k)
        p = torch.rand(B, filled.shape[1], self.V, generator=g) + 0.05
        probs.scatter_(1, filled.unsqueeze(-1).expand(-1, -1, self.V), p / p.sum(-1, keepdim=True))
        return x, ctx_out, tgt_out, [], [], probs
