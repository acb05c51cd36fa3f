"""Memory-bound kernels (K1, LayerNorm, K5-K10) and the attention kernel (K3) against the CPU oracle /
torch fp32 references on the same seeded inputs, plus the reference-generated golden fixtures."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _ops():
    from mebt_b200 import ops
    return ops


@pytest.mark.parametrize("B,NC,NT,L,D", [(2, 300, 724, 256, 256), (3, 0, 1024, 64, 128), (1, 1023, 1, 256, 1024)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embed_gather(B, NC, NT, L, D, dtype):
    ops = _ops()
    from oracle import mebt_oracle as O
    g = torch.Generator().manual_seed(1)
    V, N = 16384, NC + NT
    P = {"tok_emb.weight": torch.randn(V, D, generator=g), "pos_emb": torch.randn(1, N, D, generator=g),
         "mask_emb": torch.randn(1, 1, D, generator=g), "sos_emb": torch.randn(1, L, D, generator=g)}
    x = torch.randint(0, V, (B, N), generator=g)
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    ctx_idx, tgt_idx = perm[:, :NC], perm[:, NC:]
    lat_r, ctx_r, tgt_r = O.stem(P, {}, x, ctx_idx, tgt_idx)
    dev = {k: v.cuda() for k, v in P.items()}
    perm_d = perm.cuda()
    ctx, tgt, lat = ops.embed_gather(x.cuda(), perm_d[:, :NC], perm_d[:, NC:], dev["tok_emb.weight"], dev["pos_emb"],
                                     dev["mask_emb"], dev["sos_emb"], out_dtype=dtype)
    ops.check_index_errors()
    for got, ref in ((ctx, ctx_r), (tgt, tgt_r), (lat, lat_r.contiguous())):
        ref2 = ref.reshape(-1, D)
        if dtype == torch.float32:
            assert torch.equal(got.cpu(), ref2)              # bit-exact: one fp32 add per element
        else:
            assert torch.equal(got.cpu(), ref2.bfloat16())   # exactly the bf16 rounding of the fp32 result


def test_embed_gather_flags_bad_index():
    ops = _ops()
    from mebt_b200._lib import MebtError
    x = torch.zeros(1, 8, dtype=torch.long, device="cuda")
    idx = torch.arange(8, device="cuda").view(1, 8).clone()
    idx[0, 3] = 99
    D = 64
    ops.embed_gather(x, idx[:, :4], idx[:, 4:], torch.zeros(16, D, device="cuda"), torch.zeros(1, 8, D, device="cuda"),
                     torch.zeros(1, 1, D, device="cuda"), torch.zeros(1, 4, D, device="cuda"))
    with pytest.raises(MebtError):
        ops.check_index_errors()


@pytest.mark.parametrize("rows,D", [(1000, 256), (777, 1024), (5, 128), (64, 2048)])
@pytest.mark.parametrize("dt_in,dt_out", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32),
                                           (torch.float32, torch.bfloat16)])
def test_layernorm(rows, D, dt_in, dt_out):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.5).to(dt_in)
    w, b = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    ref = F.layer_norm(x.float(), (D,), w, b, 1e-5)
    out, mean, rstd = ops.layernorm(x.cuda(), w.cuda(), b.cuda(), out_dtype=dt_out, save_stats=True)
    if dt_out == torch.float32:
        assert (out.cpu() - ref).abs().max() < 1e-5
    else:
        assert (out.float().cpu() - ref).abs().max() < 2.5e-2       # 1 bf16 ulp at |y| ~ 4
        assert ((out.float().cpu() - ref).abs() <= ref.abs() * 2 ** -8 + 1e-6).all()
    assert (mean.cpu() - x.float().mean(-1)).abs().max() < 1e-5
    assert (rstd.cpu() - (x.float().var(-1, unbiased=False) + 1e-5).rsqrt()).abs().max() < 1e-4


def test_scatter_ids():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    B, N, NT = 4, 1024, 300
    x = torch.randint(0, 16384, (B, N), generator=g)
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    ids = torch.randint(0, 16384, (B, NT), generator=g)
    ref = x.scatter(1, perm[:, -NT:], ids)
    xd = x.cuda()
    ops.scatter_ids(xd, perm.cuda()[:, -NT:], ids.cuda())
    ops.check_index_errors()
    assert torch.equal(xd.cpu(), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_masked_ce(dtype, smoothing):
    ops = _ops()
    from oracle import mebt_oracle as O
    g = torch.Generator().manual_seed(4)
    rows, V = 700, 16384
    logits = (2.0 * torch.randn(rows, V, generator=g)).to(dtype)
    tg = torch.randint(0, V, (rows,), generator=g)
    tg[:50] = logits[:50].float().argmax(-1)                 # some top-1 hits
    top5 = logits[50:90].float().topk(5).indices
    tg[50:90] = top5[:, 3]                                    # some top-5 hits
    ce, n1, n5 = O.masked_ce(logits.float(), tg, smoothing)
    lg = logits.float().clone().requires_grad_(True)
    F.cross_entropy(lg, tg, reduction="sum", label_smoothing=smoothing).backward()
    d_logits = torch.empty(rows, V, device="cuda", dtype=dtype)
    stats, row_loss = ops.masked_ce(logits.cuda(), tg.cuda(), smoothing, dlogits=d_logits, grad_scale=0.5)
    stats = stats.cpu()
    assert abs(stats[0].item() - ce.item()) < 1e-5 * abs(ce.item())
    assert int(stats[1]) == n1 and int(stats[2]) == n5
    ref_rows = F.cross_entropy(logits.float(), tg, reduction="none", label_smoothing=smoothing)
    assert (row_loss.cpu() - ref_rows).abs().max() < 2e-5
    tol = 5e-6 if dtype == torch.float32 else 4e-3   # expf ulps on the (softmax - 0.9) entry
    assert (d_logits.float().cpu() - 0.5 * lg.grad).abs().max() < tol


def test_sample_logits_vs_reference_golden():
    """ids/scores for the exact logits + Exp(1) noise the unmodified reference consumed (fixture)."""
    ops = _ops()
    z, _ = load_golden("sample_from_logits")
    g = torch.Generator().manual_seed(int(z["seed"]))
    logits = 3.0 * torch.randn(3, 40, 16384, generator=g)
    for tag, (T, k, tp) in dict(plain=(1.0, None, None), temp=(0.7, None, None), topk=(1.0, 32, None),
                                topp=(0.9, None, 0.8), both=(0.8, 100, 0.9)).items():
        torch.manual_seed(123)
        q = torch.empty_like(logits).exponential_()
        ids, scores, probs = ops.sample_logits(logits.view(-1, 16384).cuda(), T, k, tp, noise=q.view(-1, 16384).cuda(),
                                               return_probs=True)
        ids, scores, probs = ids.cpu().view(3, 40), scores.cpu().view(3, 40), probs.cpu().view(3, 40, -1)
        nnz_diff = np.abs((probs > 0).sum(-1).numpy() - z[f"{tag}_nnz"])
        same_set = torch.from_numpy(nnz_diff == 0)
        if tp is None:
            assert (nnz_diff == 0).all()
        else:   # nucleus boundary: fp32 summation order may move it by one token in a few rows
            assert nnz_diff.max() <= 1 and (nnz_diff > 0).mean() < 0.1, (tag, nnz_diff.max(), (nnz_diff > 0).mean())
            assert torch.allclose(probs.sum(-1), torch.ones(3, 40), atol=1e-5)
        ref_ids = torch.from_numpy(z[f"{tag}_ids"])
        mism = (ids != ref_ids) & same_set
        # bit-exact selection; a differing id is admitted only as a float near-tie of the race, proven in float64: the
        # two candidates' keys p_i / q_i (p = softmax of the scaled logits) agree to NEAR_TIE relative - what separates
        # the GPU's expf from the CPU's is a few fp32 ulps (1.2e-7 each)
        NEAR_TIE = 1e-5
        if mism.any():
            p64 = torch.softmax(logits.double() / (T + 1e-8), -1)
            k64 = p64 / q.double()
            kg = k64.gather(-1, ids.unsqueeze(-1)).squeeze(-1)[mism]
            kr = k64.gather(-1, ref_ids.unsqueeze(-1)).squeeze(-1)[mism]
            assert ((kg / kr - 1).abs() < NEAR_TIE).all(), (tag, (kg / kr - 1).abs().max())
        assert mism.float().mean() <= 0.01, (tag, int(mism.sum()))
        ok = ~(ids != ref_ids) & same_set
        np.testing.assert_allclose(scores[ok].numpy(), z[f"{tag}_score"][ok.numpy()], rtol=2e-5, atol=1e-12)
        sub_same = same_set[:, ::5]
        np.testing.assert_allclose(probs[:, ::5, ::211][sub_same].numpy(), z[f"{tag}_psub"][sub_same.numpy()],
                                   rtol=2e-5 if tp is not None else 5e-6, atol=1e-12)


def test_sample_logits_bf16_and_philox_distribution():
    ops = _ops()
    V = 16384
    g = torch.Generator().manual_seed(6)
    row = torch.full((V,), -30.0)
    row[:8] = torch.tensor([2.0, 1.0, 0.0, -1.0, 1.5, 0.5, -0.5, -2.0])
    p = F.softmax(row, -1)[:8]
    rows = 20000
    logits = row.repeat(rows, 1).cuda()
    ids, scores, _ = ops.sample_logits(logits, 1.0, None, None, noise=None, seed=1234, offset=0)
    counts = torch.bincount(ids.cpu(), minlength=V)[:8].float()
    assert counts.sum() >= rows - 2
    emp = counts / rows
    assert (emp - p).abs().max() < 4 * (p * (1 - p) / rows).sqrt().max() + 1e-3
    ids2, _, _ = ops.sample_logits(logits, 1.0, None, None, noise=None, seed=1234, offset=1)
    assert (ids2 != ids).any()                                   # a different offset is a different draw
    ids3, _, _ = ops.sample_logits(logits, 1.0, None, None, noise=None, seed=1234, offset=0)
    assert torch.equal(ids3, ids)                                # same (seed, offset) is reproducible
    idsb, _, _ = ops.sample_logits(logits.bfloat16(), 1.0, 4, None, noise=None, seed=5, offset=0)
    assert set(idsb.cpu().tolist()) <= {0, 1, 4, 5}              # top-4 of the row


def test_remask_sort_vs_reference_golden():
    ops = _ops()
    z, _ = load_golden("maskgen")
    ctx, tgt, score = (torch.from_numpy(z[k]) for k in ("gnm_ctx", "gnm_tgt", "gnm_score"))
    torch.manual_seed(77)
    q = torch.empty_like(score).exponential_()
    n_new = z["gnm_next_ctx"].shape[1] - ctx.shape[1]
    nc, nt, order = ops.remask_sort(score.cuda(), ctx.cuda(), tgt.cuda(), n_new, 2.25, noise=q.cuda(), want_order=True)
    nc, nt = nc.cpu().numpy(), nt.cpu().numpy()
    # powf on the GPU vs torch.pow on the CPU differ by ulps: near-equal keys may swap.  Every position where the
    # selection differs from the reference's is proven a near-tie in float64: the keys (s / sum s) / q^ctemp of the two
    # tokens agree to NEAR_TIE relative.
    NEAR_TIE = 1e-5
    key64 = (score.double() / score.double().sum(-1, keepdim=True)) / q.double() ** 2.25
    n_ctx0 = ctx.shape[1]
    for got, ref in ((nc[:, n_ctx0:], z["gnm_next_ctx"][:, n_ctx0:]), (nt, z["gnm_next_tgt"])):
        for b in range(got.shape[0]):
            of = {int(t): float(k) for t, k in zip(tgt[b].tolist(), key64[b].tolist())}
            for a, r in zip(got[b][got[b] != ref[b]].tolist(), ref[b][got[b] != ref[b]].tolist()):
                assert abs(of[a] / of[r] - 1.0) < NEAR_TIE, (a, r, of[a], of[r])
    assert np.array_equal(nc[:, :n_ctx0], z["gnm_next_ctx"][:, :n_ctx0])
    assert (nc == z["gnm_next_ctx"]).mean() > 0.99 and (nt == z["gnm_next_tgt"]).mean() > 0.99
    assert sorted(np.concatenate([nc[0], nt[0]]).tolist()) == list(range(1024))


@pytest.mark.parametrize("NT,ctemp", [(824, 0.0), (8192, 3.0), (1, 1.0), (4097, 0.5)])
def test_remask_sort_order_property(NT, ctemp):
    ops = _ops()
    from oracle import mebt_oracle as O
    g = torch.Generator().manual_seed(8)
    B, NC = 3, 17
    score = torch.rand(B, NT, generator=g)
    q = torch.empty(B, NT).exponential_(generator=g)
    tgt = torch.stack([torch.randperm(NT + NC, generator=g) for _ in range(B)])
    ctx, tgt = tgt[:, :NC], tgt[:, NC:]
    n_new = NT // 3
    nc, nt, order = ops.remask_sort(score.cuda(), ctx.cuda(), tgt.cuda(), n_new, ctemp, noise=q.cuda(), want_order=True)
    order = order.cpu()
    ref = O.remask_order(score, ctemp, q)
    key = (score / score.sum(-1, keepdim=True)) / (q ** ctemp)
    sorted_keys = key.gather(1, order)
    assert (sorted_keys[:, 1:] <= sorted_keys[:, :-1] * (1 + 1e-5)).all()       # descending up to pow ulps
    assert (order.sort(1).values == torch.arange(NT)).all()                      # a permutation
    assert (order == ref).float().mean() > 0.99
    # every differing position holds two tokens whose float64 keys agree to 1e-5 relative (a documented near-tie)
    key64 = (score.double() / score.double().sum(-1, keepdim=True)) / q.double() ** ctemp
    diff = order != ref
    if diff.any():
        kg, kr = key64.gather(1, order)[diff], key64.gather(1, ref)[diff]
        assert ((kg / kr - 1).abs() < 1e-5).all(), (kg / kr - 1).abs().max()
    if ctemp == 0.0:
        assert torch.equal(order, ref)                                            # no pow involved: bit-exact
    assert torch.equal(nc.cpu(), torch.cat([ctx, tgt.gather(1, order[:, :n_new])], 1))
    assert torch.equal(nt.cpu(), tgt.gather(1, order[:, n_new:]))


@pytest.mark.parametrize("tensor_cores", [True, False])
def test_vq_argmin_and_gather_vs_reference_golden(tensor_cores):
    """Both search kernels against the reference's recorded encodings: the fp16-split tcgen05 GEMM with the argmin epilogue
    (the default) and the fp32 FFMA kernel, under the same near-tie rule."""
    ops = _ops()
    z, _ = load_golden("codebook")
    torch.manual_seed(int(z["cb_seed"]))
    E = torch.randn(16384, 256)
    g = torch.Generator().manual_seed(int(z["z_seed"]))
    zz = torch.randn(2, 256, 4, 16, 16, generator=g)
    enc = ops.vq_argmin(zz.cuda(), E.cuda(), tensor_cores=tensor_cores)
    ref = torch.from_numpy(z["encodings"])
    mism = (enc.cpu() != ref)
    gap = torch.from_numpy(z["gap"]).view_as(ref)
    # bit-exact indices; a flip is only admissible where best and runner-up differ by < 1e-3 (fp32 ulp at
    # |d| ~ 370 is 3e-5 and the dot-product summation order differs from MKL's)
    assert (gap[mism] < 1e-3).all(), gap[mism]
    assert mism.sum() <= 2
    emb = ops.row_gather(ref.cuda(), E.cuda(), channel_first=True)
    assert torch.equal(emb.cpu(), F.embedding(ref, E).permute(0, 4, 1, 2, 3).contiguous())
    emb2 = ops.row_gather(ref.cuda(), E.cuda(), channel_first=False)
    assert torch.equal(emb2.cpu(), F.embedding(ref, E))
    ops.check_index_errors()


def test_vq_argmin_tensor_cores_vs_float64():
    """The split arithmetic on its own bar: against float64 distances every chosen code is the true nearest one or within
    1e-4 of it (the fp32 rounding of d ~ 370 alone is 3e-5), on ragged sizes (rows not a multiple of the tile, a batch
    boundary inside a tile) and with exact duplicates in the codebook (lowest index wins)."""
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    E = torch.randn(4096, 128, generator=g)
    E[77] = E[5]                                              # exact tie: index 5 must win over 77
    z = torch.randn(3, 128, 1, 7, 13, generator=g)           # 91 vectors per batch element
    z[0, :, 0, 0, 0] = E[77]                                  # a vector that sits exactly on the duplicated code
    enc = ops.vq_argmin(z.cuda(), E.cuda(), tensor_cores=True).cpu()
    flat = z.permute(0, 2, 3, 4, 1).reshape(-1, 128).double()
    d = (flat ** 2).sum(1, keepdim=True) - 2 * flat @ E.double().t() + (E.double() ** 2).sum(1)[None]
    best = d.min(1).values
    chosen = d.gather(1, enc.reshape(-1, 1)).squeeze(1)
    assert (chosen - best).max() < 1e-4, (chosen - best).max()
    assert enc.reshape(-1)[0] == 5
    enc_f = ops.vq_argmin(z.cuda(), E.cuda(), tensor_cores=False).cpu()
    assert (enc != enc_f).sum() <= 1


def _attn_ref(q, k, v):
    s = (q.float() @ k.float().transpose(-1, -2)) * 0.125
    return F.softmax(s, -1) @ v.float()


@pytest.mark.parametrize("B,H,NQ,NK1,NK2", [(2, 4, 256, 300, 0), (1, 2, 256, 256, 0), (2, 2, 200, 256, 0),
                                             (2, 4, 256, 256, 724), (1, 16, 256, 1, 0), (2, 2, 256, 0, 0),
                                             (1, 2, 128, 0, 130), (1, 1, 1, 4096, 0)])
def test_latent_attention(B, H, NQ, NK1, NK2):
    ops = _ops()
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(9)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
    qbuf = rnd(B * NQ, 3 * D)                     # Q lives at column offset D of a wider buffer
    kv1 = rnd(B * NK1, 2 * D) if NK1 else None
    kv2 = rnd(B * NK2, 2 * D) if NK2 else None
    lse = torch.empty(B, H, NQ, device="cuda")
    out = ops.attention(qbuf, D, kv1, 0, D, NK1, kv2, D, 0, NK2, B, H, NQ, lse=lse)
    torch.cuda.synchronize()
    q = qbuf[:, D:2 * D].view(B, NQ, H, 64).transpose(1, 2)
    ks, vs = [], []
    if NK1:
        ks.append(kv1[:, :D].view(B, NK1, H, 64).transpose(1, 2))
        vs.append(kv1[:, D:].view(B, NK1, H, 64).transpose(1, 2))
    if NK2:
        ks.append(kv2[:, D:].view(B, NK2, H, 64).transpose(1, 2))
        vs.append(kv2[:, :D].view(B, NK2, H, 64).transpose(1, 2))
    if not ks:
        assert (out == 0).all()
        return
    k, v = torch.cat(ks, 2), torch.cat(vs, 2)
    ref = _attn_ref(q, k, v).transpose(1, 2).reshape(B * NQ, D)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err
    s = (q.float() @ k.float().transpose(-1, -2)) * 0.125
    assert (lse - torch.logsumexp(s, -1)).abs().max() < 1e-3


@pytest.mark.parametrize("B,H,NQ,NK1,NK2", [(1, 16, 256, 8192, 0), (2, 16, 256, 256, 8192), (1, 2, 200, 1500, 0),
                                             (2, 4, 256, 4096 - 37, 0)])
def test_latent_attention_split_kv(B, H, NQ, NK1, NK2):
    """Small-batch sampling shapes (few work items, thousands of keys) take the split-KV form: each item's key list is
    divided over several CTAs and the partial rows are merged by a second kernel.  Same bar as the unsplit kernel against
    the fp32 reference, and agreement with the unsplit kernel itself (requesting the log-sum-exp disables the split)."""
    ops = _ops()
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(21)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
    qbuf = rnd(B * NQ, D)
    kv1 = rnd(B * NK1, 2 * D) if NK1 else None
    kv2 = rnd(B * NK2, 2 * D) if NK2 else None
    out = ops.attention(qbuf, 0, kv1, 0, D, NK1, kv2, 0, D, NK2, B, H, NQ)
    lse = torch.empty(B, H, NQ, device="cuda")
    unsplit = ops.attention(qbuf, 0, kv1, 0, D, NK1, kv2, 0, D, NK2, B, H, NQ, lse=lse)
    torch.cuda.synchronize()
    q = qbuf.view(B, NQ, H, 64).transpose(1, 2)
    ks = [t[:, :D].view(B, -1, H, 64).transpose(1, 2) for t in (kv1, kv2) if t is not None]
    vs = [t[:, D:].view(B, -1, H, 64).transpose(1, 2) for t in (kv1, kv2) if t is not None]
    ref = _attn_ref(q, torch.cat(ks, 2), torch.cat(vs, 2)).transpose(1, 2).reshape(B * NQ, D)
    scale = max(1.0, ref.abs().max().item())
    assert (out.float() - ref).abs().max().item() < 2e-2 * scale
    assert (out.float() - unsplit.float()).abs().max().item() < 1e-2 * scale


def test_head_sample_fused_gumbel_max():
    """The sampling step fused into the head GEMM (no logits in HBM): draws follow softmax(logits / T) - checked on 60000
    rows that share one logit vector, category by category within 5 sigma plus the total-variation distance -, differ
    between (seed, offset) pairs and are reproducible for the same pair; a second, peaked row type checks the temperature."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    D, V, rows = 64, 1024, 60000
    w = (0.35 * torch.randn(V, D, generator=g)).bfloat16()
    x1 = torch.randn(D, generator=g).bfloat16()
    for T in (1.0, 0.7):
        x = x1.repeat(rows, 1).contiguous()
        ids = ops.head_sample(x.cuda(), w.cuda(), T, seed=11, offset=3).cpu()
        assert int(ids.min()) >= 0 and int(ids.max()) < V
        logits = (w.float() @ x1.float()).double()                     # what the tensor cores accumulate (fp32) from bf16 operands
        p = torch.softmax(logits / (T + 1e-8), 0)
        emp = torch.bincount(ids, minlength=V).double() / rows
        sigma = (p * (1 - p) / rows).sqrt()
        assert ((emp - p).abs() <= 5 * sigma + 1e-4).all(), ((emp - p).abs() / (sigma + 1e-12)).max()
        assert 0.5 * (emp - p).abs().sum() < 0.06                      # total variation (sampling noise alone gives ~0.05)
    x = x1.repeat(rows, 1).contiguous().cuda()
    a = ops.head_sample(x, w.cuda(), 1.0, seed=11, offset=3)
    b = ops.head_sample(x, w.cuda(), 1.0, seed=11, offset=3)
    c = ops.head_sample(x, w.cuda(), 1.0, seed=11, offset=4)
    d = ops.head_sample(x, w.cuda(), 1.0, seed=12, offset=3)
    assert torch.equal(a, b) and (a != c).float().mean() > 0.5 and (a != d).float().mean() > 0.5
    # rows draw independently: neighbouring rows agree no more often than sum_v p_v^2 predicts
    same = (a[1:] == a[:-1]).float().mean().item()
    logits = (w.float() @ x1.float()).double()
    p = torch.softmax(logits, 0)
    assert abs(same - float((p * p).sum())) < 0.01
