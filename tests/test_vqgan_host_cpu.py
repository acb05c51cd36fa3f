"""Host logic of the VQGAN convolution wrappers (mebt_b200/vqgan.py) without a GPU: the packed operands and the launch
arguments they hand to `mebt_conv3d_ndhwc` are run through a plain-torch restatement of that entry point's CONTRACT
(include/mebt_b200.h: y[b, t*ystep+yorigin, ...] = bias + sum_{taps, c} xp[b, t*step + origin + dt, ...] * w[co][(tap)][c],
window mode when cin > ldx) and compared with torch.nn.functional on the reference's formulation
(mebt/vqgan.py:358-405: F.pad(replicate) + Conv3d / ConvTranspose3d(padding = k - 1))."""
import itertools

import pytest
import torch
import torch.nn.functional as F

from mebt_b200 import vqgan as V


def conv3d_ndhwc_contract(xp, w, cin, cout, taps, step, odims, bias, out=None, origin=(0, 0, 0), ystep=(1, 1, 1), yorigin=(0, 0, 0)):
    """xp [B, Tp, Hp, Wp, ld] fp32, w [rows >= cout, prod(taps) * ceil64(cin)]; -> channels-last fp32 output."""
    B, Tp, Hp, Wp, ld = xp.shape
    window = cin > ld
    cp = -(-cin // 64) * 64
    assert w.shape[0] % 64 == 0 and w.shape[0] >= cout and w.shape[1] == taps[0] * taps[1] * taps[2] * cp
    wf = w.float().view(w.shape[0], *taps, cp)
    assert float(wf[cout:].abs().max() if w.shape[0] > cout else 0.0) == 0.0          # zero rows behind cout
    To, Ho, Wo = odims
    y = torch.zeros(B, To, Ho, Wo, cout)
    rows = xp.reshape(B, Tp, Hp, Wp * ld)
    for dt, dh, dw in itertools.product(*(range(k) for k in taps)):
        ts = origin[0] + dt + step[0] * torch.arange(To)
        hs = origin[1] + dh + step[1] * torch.arange(Ho)
        ws = origin[2] + dw + step[2] * torch.arange(Wo)
        if window:          # the K slice of a position: the cin consecutive ELEMENTS from its first channel on
            assert taps[2] == 1 and step[2] == 1 and int(ws[-1]) * ld + cin <= Wp * ld
            a = torch.stack([rows[:, ts][:, :, hs][..., int(w0) * ld:int(w0) * ld + cin] for w0 in ws], dim=3)
        else:
            a = xp[:, ts][:, :, hs][:, :, :, ws][..., :cin]
        y += a @ wf[:cout, dt, dh, dw, :cin].t()
    y += bias[:cout]
    if out is None:
        return y
    out[:, yorigin[0]::ystep[0], yorigin[1]::ystep[1], yorigin[2]::ystep[2], :cout] = y
    return out


def channels_last(x, ld):
    out = torch.zeros(x.shape[0], *x.shape[2:], ld)
    out[..., :x.shape[1]] = x.permute(0, 2, 3, 4, 1)
    return out


def pad_cl(x, pads6):       # replicate padding of a channels-last tensor, pads6 = (t0, t1, h0, h1, w0, w1)
    y = F.pad(x.permute(0, 4, 1, 2, 3), (pads6[4], pads6[5], pads6[2], pads6[3], pads6[0], pads6[1]), mode="replicate")
    return y.permute(0, 2, 3, 4, 1).contiguous()


@pytest.mark.parametrize("cin,cout,k,stride,dims", [(16, 24, 3, (1, 1, 1), (3, 5, 6)), (8, 40, 4, (2, 2, 2), (4, 6, 8)),
                                                    (70, 8, 4, (1, 2, 2), (3, 4, 6)), (5, 3, 1, (1, 1, 1), (2, 3, 4))])
def test_same_pad_conv3d_packing(cin, cout, k, stride, dims):
    torch.manual_seed(0)
    m = V.SamePadConv3d(cin, cout, k, stride=stride)
    x = torch.randn(2, cin, *dims)
    ref = F.conv3d(F.pad(x, m.pad_input, mode="replicate"), m.conv.weight.to(torch.bfloat16).float(), m.conv.bias, stride=stride)
    w, b = m._operands()
    pads = m.pads
    xp = pad_cl(channels_last(x, -(-cin // 8) * 8), (pads[0][0], pads[0][1], pads[1][0], pads[1][1], pads[2][0], pads[2][1]))
    odims = tuple(d // s for d, s in zip(dims, stride))
    got = conv3d_ndhwc_contract(xp, w, cin, b.shape[0], m.kernel_size, m.stride, odims, b)
    assert b.shape[0] == -(-cout // 8) * 8 and w.shape[0] == -(-cout // 64) * 64
    assert torch.allclose(got[..., :cout].permute(0, 4, 1, 2, 3), ref, atol=2e-4, rtol=1e-4)
    assert float(got[..., cout:].abs().max() if got.shape[-1] > cout else 0.0) == 0.0


@pytest.mark.parametrize("cin,kw", [(3, 3), (8, 3), (2, 5)])
def test_window_mode_packing_of_a_few_channel_input(cin, kw):
    """The 3-channel first convolution: the taps along w packed into one 64-element K slice (8 positions x 8 channels);
    the row padded by 8 - kw extra positions on the right (`SamePadConv3d._window` / `_operands_window`)."""
    torch.manual_seed(1)
    m = V.SamePadConv3d(cin, 32, (3, 3, kw))
    x = torch.randn(1, cin, 2, 3, 128)
    xcl = channels_last(x, 8)
    assert m._window(xcl)
    ref = F.conv3d(F.pad(x, m.pad_input, mode="replicate"), m.conv.weight.to(torch.bfloat16).float(), m.conv.bias)
    w, b = m._operands_window()
    pads = m.pads
    xp = pad_cl(xcl, (pads[0][0], pads[0][1], pads[1][0], pads[1][1], pads[2][0], pads[2][1] + 8 - kw))
    got = conv3d_ndhwc_contract(xp, w, 64, b.shape[0], (3, 3, 1), (1, 1, 1), (2, 3, 128), b)
    assert w.shape == (64, 9 * 64)
    assert torch.allclose(got.permute(0, 4, 1, 2, 3), ref, atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
def test_transposed_convolution_as_parity_classes(stride):
    """ConvTranspose3d(kernel 4, stride s, padding 3) over the replicate-padded input = one stride-1 convolution per output
    parity with the matching kernel slices, written to the interleaved positions (vqgan.py:384-405)."""
    torch.manual_seed(2)
    m = V.SamePadConvTranspose3d(12, 20, 4, stride=stride)
    x = torch.randn(2, 12, 3, 4, 5)
    ref = F.conv_transpose3d(F.pad(x, m.pad_input, mode="replicate"), m.convt.weight.to(torch.bfloat16).float(), m.convt.bias,
                             stride=stride, padding=3)
    classes, b = m._operands()
    assert len(classes) == 2 ** sum(s == 2 for s in stride)
    pads = m.pads
    xp = pad_cl(channels_last(x, 16), (pads[0][0], pads[0][1], pads[1][0], pads[1][1], pads[2][0], pads[2][1]))
    odims = (3, 4, 5)
    out = torch.zeros(2, *(d * s for d, s in zip(odims, stride)), b.shape[0])
    for par, taps, w in classes:
        conv3d_ndhwc_contract(xp, w, 12, b.shape[0], taps, (1, 1, 1), odims, b, out=out, origin=par, ystep=stride, yorigin=par)
    assert ref.shape[2:] == out.shape[1:4]
    assert torch.allclose(out[..., :20].permute(0, 4, 1, 2, 3), ref, atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("tag", ["small", "bn"])
def test_vqgan_state_dict_names_and_shapes_equal_the_reference(tag):
    """Drop-in contract of the checkpoint format: every encoder / decoder / pre- / post-quant parameter and buffer of the
    unmodified reference's modules (names and shapes recorded by tests/golden/make_golden.py::gen_vqgan) exists in
    `mebt_b200.vqgan.VQGAN` under the same name with the same shape, and nothing else does (codebook buffers aside) - so a
    reference checkpoint loads with strict key matching."""
    import json
    from conftest import load_golden
    z, _ = load_golden(f"vqgan_{tag}")
    cfg = json.loads(str(z["cfg_json"]))
    want = {k: tuple(v) for k, v in json.loads(str(z["shapes_json"])).items()}
    cfg.update(sequence_length=8, sample_every_n_frames=1, resolution=32)
    model = V.VQGAN(V._Args(cfg))
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    mine = {k: s for k, s in got.items() if not k.startswith("codebook.") or k == "codebook.embeddings"}
    assert mine == want, (sorted(set(mine) ^ set(want))[:8], [k for k in mine if k in want and mine[k] != want[k]][:8])
    assert {k for k in got if k.startswith("codebook.")} == {"codebook.embeddings", "codebook.N", "codebook.z_avg"}


def test_vqgan_command_line_hyperparameters():
    """`VQGAN.add_model_specific_args` (vqgan.py:229-252): the flags, types and defaults of the reference's command line, so
    that its training / sampling scripts parse unchanged and a default namespace builds the default model."""
    import argparse
    parser = V.VQGAN.add_model_specific_args(argparse.ArgumentParser(add_help=False))
    d = vars(parser.parse_args([]))
    assert d == dict(embedding_dim=256, n_codes=2048, n_hiddens=240, lr=3e-4, downsample=(4, 4, 4), disc_channels=64, disc_layers=3,
                     discriminator_iter_start=50000, disc_loss_type="hinge", image_gan_weight=1.0, video_gan_weight=1.0,
                     l1_weight=4.0, gan_feat_weight=0.0, perceptual_weight=0.0, i3d_feat=False, restart_thres=1.0,
                     no_random_restart=False, norm_type="group", padding_type="replicate")
    a = parser.parse_args(["--n_codes", "16384", "--n_hiddens", "32", "--downsample", "4", "8", "8", "--norm_type", "batch"])
    assert a.downsample == [4, 8, 8] and a.norm_type == "batch"
    with pytest.raises(SystemExit):
        parser.parse_args(["--disc_loss_type", "wasserstein"])
    model = V.VQGAN(args=a)                                   # the reference's only constructor argument is `args`
    assert isinstance(model.encoder, V.Encoder) and model.codebook.n_codes == 16384
    assert isinstance(model.encoder.final_block[0], torch.nn.BatchNorm3d)
    assert float(V.silu(torch.tensor(0.0))) == 0.0 and abs(float(V.silu(torch.tensor(1.0))) - 0.7310586) < 1e-6
