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
