"""Teacher-forced parity of the sampler LOOPS on the GPU — `Net2NetTransformer.sample` (maskgit / random / bootstrap,
mebt/transformer.py:353-447 with MaskGen.generate_next_mask, mask_sampler.py:189-246) and `entp_sample`
(transformer.py:449-542 with generate_next_mask_entp, mask_sampler.py:248-303) — against the CPU oracle, which is
itself pinned bit-for-bit to the unmodified reference (tests/golden/sampling_*.npz, entp_micro.npz).

A bf16 GPU forward and the fp32 CPU reference legitimately differ by up to 1e-2 in the logits, which flips Gumbel-max
winners, so end-to-end ids are not comparable.  The loops are therefore run for real on the CUDA model with
  * the oracle's noise draws (rng hook: same torch CPU generator, same order),
  * each step's GPU logits CHECKED against the oracle's logits for the same state and then replaced by them,
  * each step's GPU confidence scores CHECKED against the oracle's and replaced by them before re-masking,
and everything the loop itself is responsible for — step skipping, the float32 mask-size arithmetic, the fused
sampling kernel's selections, the write-back, the re-mask sort and the index bookkeeping — must then reproduce the
oracle's final ids / context / target index tensors BIT FOR BIT."""
import pytest
import torch

from conftest import load_golden
from helpers import build_model

pytestmark = pytest.mark.gpu
TOL = 1e-2
V = 16384


class Teacher:
    def __init__(self, model, P, cfg, temperature, top_k, top_p, entropy):
        from oracle import mebt_oracle as O
        self.O, self.model, self.P, self.cfg = O, model, P, cfg
        self.T, self.k, self.p, self.entropy = temperature, top_k, top_p, entropy
        self.ref_logits = self.q = self.q_mask = None
        self.steps = 0
        self.worst_logit = self.worst_score = 0.0
        self.tie_swaps = 0
        self.randn = None

    # rng hook: the draws the oracle consumes, from the same CPU generator in the same order
    def draw(self, kind, shape, device):
        if kind == "exponential":
            t = torch.empty(shape, dtype=torch.float32).exponential_()
            if shape[-1] == V:
                self.q = t
            else:
                self.q_mask = t
            return t.to(device)
        self.randn = torch.randn(shape)
        return self.randn.to(device)

    def install(self):
        from mebt_b200 import rng
        model, O = self.model, self.O
        rng.set_hook(self.draw)
        real_logits = model._logits_rows
        ms = model.mask_sampler
        real_next, real_next_entp = ms.generate_next_mask, ms.generate_next_mask_entp

        def logits_rows(partial, ctx, tgt, logits_dtype=None):
            got = real_logits(partial, ctx, tgt, torch.float32)
            ref = O.reconstruct_mask(self.P, self.cfg, partial.cpu(), ctx.cpu(), tgt.cpu())
            err = (got.cpu().view_as(ref) - ref).abs().max().item() / ref.abs().max().item()
            self.worst_logit = max(self.worst_logit, err)
            assert err < TOL, f"step {self.steps}: logits off by {err}"
            self.ref_logits = ref
            self.steps += 1
            return ref.reshape(-1, V).cuda()

        def ref_scores():
            ids, probs = O.sample_from_logits(self.ref_logits, self.T, self.k, self.p, self.q.view_as(self.ref_logits))
            if self.entropy:
                return O.entropy_scores(probs)
            return probs.gather(-1, ids.unsqueeze(-1)).squeeze(-1)

        def check(score):
            ref = ref_scores()
            if self.entropy:
                # max_row(s) - s with s = sum over 16384 entries of (p - log(p + 1e-8)) ~ 3e5: the difference of two such
                # sums carries their absolute rounding error, so the bar is absolute: 3e-6 of the summands' magnitude
                err = (score.cpu() - ref).abs().max().item()
                self.worst_score = max(self.worst_score, err)
                assert err < 1.0, f"entropy scores off by {err}"
            else:
                rel = ((score.cpu() - ref).abs() / (ref.abs() + 1e-12)).max().item()
                self.worst_score = max(self.worst_score, rel)
                assert rel < 2e-5, f"scores off by {rel}"
            return ref.to(score.device)

        def next_mask(ctx, tgt, score, *a, **k):
            return real_next(ctx, tgt, check(score), *a, **k)

        def next_mask_entp(ctx, tgt, score, *a, **k):
            ref = check(score)
            self.q_mask = self.randn = None
            out = real_next_entp(ctx, tgt, ref, *a, **k)
            if self.q_mask is None:
                return out                       # nothing was revealed at this step
            # The re-mask step in isolation: same scores, same noise.  Entropy scores are differences of two fp32 sums
            # of magnitude 3e5, i.e. multiples of 2^-5: rows hold many EXACTLY equal keys, and torch.sort (not stable
            # unless asked) orders equal keys arbitrarily (CPU and CUDA torch disagree with each other as well); the
            # kernel orders them by ascending index.  So the bar is: position by position the GPU order and the oracle
            # order carry bit-identical keys, and both are permutations of the targets - they may differ only inside
            # groups of exactly tied keys (counted in `tie_swaps`).  The loop then continues from the oracle's order.
            NC = ctx.shape[1]
            n_new = out[0].shape[1] - NC
            used = self.randn if k.get("strategy", "maskgit") == "random" else ref.cpu()
            want = O.generate_next_mask(ctx.cpu(), tgt.cpu(), used, tgt.shape[1] - n_new, 0.0, self.q_mask)
            assert torch.equal(out[0][:, :NC].cpu(), want[0][:, :NC])
            got_seq = torch.cat([out[0][:, NC:], out[1]], 1).cpu()
            want_seq = torch.cat([want[0][:, NC:], want[1]], 1)
            assert torch.equal(got_seq.sort(1)[0], tgt.cpu().sort(1)[0]), "GPU re-mask order is not a permutation"
            keys = used / used.sum(-1, keepdim=True)
            by_token = torch.zeros(ctx.shape[0], int(tgt.max()) + 1).scatter_(1, tgt.cpu(), keys)
            assert torch.equal(by_token.gather(1, got_seq), by_token.gather(1, want_seq)), \
                f"re-mask order differs beyond exact ties at loop step {self.steps}"
            self.tie_swaps += int((got_seq != want_seq).sum())
            return want[0].to(ctx.device), want[1].to(ctx.device)

        model._logits_rows = logits_rows
        # entp_sample reaches generate_next_mask through generate_next_mask_entp: check the scores only once
        if self.entropy:
            ms.generate_next_mask_entp = next_mask_entp
        else:
            ms.generate_next_mask = next_mask
        self._undo = (real_logits, real_next, real_next_entp)

    def remove(self):
        from mebt_b200 import rng
        rng.set_hook(None)
        self.model._logits_rows, self.model.mask_sampler.generate_next_mask, self.model.mask_sampler.generate_next_mask_entp = self._undo


@pytest.fixture(scope="module")
def micro():
    from oracle import mebt_oracle as O
    z, cfg = load_golden("sampling_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    return cfg, P, build_model(cfg, P, schedule="cosine")


@pytest.mark.parametrize("strategy,steps,T,k,ctemp", [("maskgit", 6, 1.0, None, 4.5), ("maskgit", 5, 0.8, 64, 6.0),
                                                      ("random", 4, 1.0, None, 4.5), ("bootstrap", 3, 1.0, None, 4.5)])
def test_sample_loop_teacher_forced(micro, strategy, steps, T, k, ctemp):
    from oracle import mebt_oracle as O
    cfg, P, model = micro
    B, N = 2, 256
    x0 = torch.zeros(B, N, dtype=torch.long)
    ref_ids, ref_ctx, ref_tgt = O.sample_maskgit(P, cfg, x0, O.TorchRng(31), temperature=T, top_k=k, n_steps=steps,
                                                 strategy=strategy, context_temperature=ctemp, schedule_name="cosine")
    teacher = Teacher(model, P, cfg, T, k, None, entropy=False)
    teacher.install()
    try:
        torch.manual_seed(31)
        ids, ctx, tgt = model.sample(x0.cuda(), None, T, k, None, n_steps=steps, strategy=strategy, context_temperature=ctemp)
    finally:
        teacher.remove()
    assert teacher.steps >= 2
    assert torch.equal(ids.cpu(), ref_ids) and torch.equal(ctx.cpu(), ref_ctx) and torch.equal(tgt.cpu(), ref_tgt)


@pytest.mark.parametrize("edit,steps,n_keep", [(True, 5, 192), (True, 4, 64), (False, 6, 128)])
def test_sample_loop_with_fixed_context_and_edit_teacher_forced(micro, edit, steps, n_keep):
    """`sample(..., context_indices, target_indices, edit=True)` - the call `extrapolate()` makes for every window of a
    long video (sample_vqgan_transformer_videos.py:95-157): the first `n_keep` tokens are given context, only the rest is
    re-sampled, and with `edit` the schedule counts masked tokens against the number of TARGETS (transformer.py:373-376).
    Teacher-forced like the loops above: every step's logits within tolerance, then ids / masks bit for bit."""
    from oracle import mebt_oracle as O
    cfg, P, model = micro
    B, N = 2, 256
    g = torch.Generator().manual_seed(7)
    x0 = torch.randint(0, cfg["vocab_size"], (B, N), generator=g)
    ctx0 = torch.stack([torch.randperm(n_keep, generator=g) for _ in range(B)])          # kept frames, in some order
    tgt0 = torch.stack([n_keep + torch.randperm(N - n_keep, generator=g) for _ in range(B)])
    ref_ids, ref_ctx, ref_tgt = O.sample_maskgit(P, cfg, x0, O.TorchRng(47), temperature=1.0, n_steps=steps, strategy="maskgit",
                                                 context_temperature=4.5, schedule_name="cosine", context_indices=ctx0,
                                                 target_indices=tgt0, edit=edit)
    teacher = Teacher(model, P, cfg, 1.0, None, None, entropy=False)
    teacher.install()
    try:
        torch.manual_seed(47)
        ids, ctx, tgt = model.sample(x0.cuda(), None, 1.0, None, None, n_steps=steps, strategy="maskgit", context_temperature=4.5,
                                     context_indices=ctx0.cuda(), target_indices=tgt0.cuda(), edit=edit)
    finally:
        teacher.remove()
    assert teacher.steps >= 2
    assert torch.equal(ids.cpu(), ref_ids) and torch.equal(ctx.cpu(), ref_ctx) and torch.equal(tgt.cpu(), ref_tgt)
    assert torch.equal(ids.cpu()[:, :n_keep], x0[:, :n_keep])                             # the given context is never rewritten
    assert torch.equal(ctx.cpu()[:, :n_keep], ctx0)


@pytest.mark.parametrize("strategy,steps", [("maskgit", 5), ("random", 4), ("bootstrap", 3)])
def test_entp_sample_loop_teacher_forced(micro, strategy, steps):
    from oracle import mebt_oracle as O
    cfg, P, model = micro
    B, N = 2, 256
    x0 = torch.zeros(B, N, dtype=torch.long)
    ref_ids, ref_ctx, ref_tgt = O.sample_entp(P, cfg, x0, O.TorchRng(13), n_steps=steps, strategy=strategy,
                                              schedule_name="cosine")
    teacher = Teacher(model, P, cfg, 1.0, None, None, entropy=True)
    teacher.install()
    try:
        torch.manual_seed(13)
        ids, ctx, tgt = model.entp_sample(x0.cuda(), None, n_steps=steps, strategy=strategy)
    finally:
        teacher.remove()
    assert teacher.steps >= 2
    assert ctx.shape == ref_ctx.shape and tgt.shape == ref_tgt.shape, (ctx.shape, ref_ctx.shape)
    assert torch.equal(ctx.cpu(), ref_ctx) and torch.equal(tgt.cpu(), ref_tgt)
    assert torch.equal(ids.cpu(), ref_ids), (ids.cpu() != ref_ids).sum()
    # ... and the oracle run with this seed is the one the reference fixture pins
    z, _ = load_golden("entp_micro")
    assert (ref_ids.numpy() == z[f"entp_{strategy}_ids"]).all()
