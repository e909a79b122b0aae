"""Algorithmic FLOP census of one network call, i.e. 2 x MACs of every conv / linear / QK^T / PV the REFERENCE executes
(what torch.utils.flop_counter.FlopCounterMode reports on the reference modules; norms, softmax, activations and
elementwise ops excluded) - the roofline numerator of SURVEY.md section 8(d).  It follows the reference's structure
(UNetModel.__init__ openaimodel.py:1254-1527, ControlNet2D controlmodel.py:196-317), NOT what the CUDA build executes:
e.g. the reference projects the text context once per frame (attention.py:1159-1163) while the build does it once per
batch entry, and the build pads 4 input channels to 8.  tests/test_host_logic.py checks it against FlopCounterMode on
the oracle and against the survey's totals (77.68 TF tv2v / 110.31 TF tvi2v at CFG batch 2 x 17 x 64 x 96).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

HINT_WIDTHS = (3, 16, 16, 32, 32, 96, 96, 256, 320)
HINT_STRIDES = (1, 1, 2, 1, 2, 1, 2, 1)


def _plan(model_channels=320, channel_mult=(1, 2, 4, 4), num_res_blocks=2, attention_resolutions=(4, 2, 1)):
    """[(kind, cin, cout, ds)] for input blocks 1.., middle, output blocks (openaimodel.py:1262-1510)."""
    mc = model_channels
    inp, chans = [], [mc]
    ch, ds = mc, 1
    for level, m in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            inp.append(("res", ch, m * mc, ds))
            ch = m * mc
            if ds in attention_resolutions:
                inp.append(("attn", ch, ch, ds))
            chans.append(ch)
        if level != len(channel_mult) - 1:
            inp.append(("down", ch, ch, ds))
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch, ds), ("attn", ch, ch, ds), ("res", ch, ch, ds)]
    out = []
    for level, m in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            out.append(("res", ch + ich, mc * m, ds))
            ch = mc * m
            if ds in attention_resolutions:
                out.append(("attn", ch, ch, ds))
            if level and i == num_res_blocks:
                out.append(("up", ch, ch, ds))
                ds //= 2
    return inp, mid, out, chans


def network_flops(kind: str, B: int, T: int, h: int, w: int, context_len: int = 77, context_dim: int = 768,
                  emb_dim: int = 1280) -> Dict[str, float]:
    """FLOPs of ONE call of OpenAIWrapperControlLDM3DTV2V.forward with x [B, 4, T, h, w] (B includes CFG doubling).
    Returns a dict by op class plus 'total'."""
    f = defaultdict(float)
    inp, mid, out, _ = _plan()

    def res(F, Bemb, L, cin, cout, T3d):          # T3d: 0 for 2-D, else T (pseudo-3D with Bemb = B)
        f["conv3x3"] += 2.0 * F * L * 9 * (cin * cout + cout * cout)
        f["linear_bias"] += 2.0 * Bemb * emb_dim * cout
        if cin != cout:
            f["conv1x1"] += 2.0 * F * L * cin * cout
        if T3d:
            f["conv1d_k3"] += 2 * 2.0 * F * L * 3 * cout * cout
            if cin != cout:
                f["conv1d_k1"] += 2.0 * F * L * cout * cout

    def single_layer(M, C, attn_flops, Mkv=None):  # BasicTransformerSingleLayerBlock: q,k,v + attention + to_out + FF
        f["linear_nobias"] += 2.0 * M * C * C + 2 * 2.0 * (M if Mkv is None else Mkv) * C * C
        f["attention"] += attn_flops
        f["linear_bias"] += 2.0 * M * C * C + 2.0 * M * C * 8 * C + 2.0 * M * 4 * C * C

    def attn(F, L, C, T3d, text=True, ca_type=None):
        M = F * L
        f["conv1x1"] += 2 * 2.0 * M * C * C                               # proj_in, proj_out
        if text:                                                          # BasicTransformerBlock
            f["linear_nobias"] += 3 * 2.0 * M * C * C                     # attn1 q, k, v
            f["attention"] += 4.0 * F * L * L * C
            f["linear_nobias"] += 2.0 * M * C * C + 2 * 2.0 * F * context_len * context_dim * C   # attn2 q; k, v per frame
            f["attention_text"] += 4.0 * F * L * context_len * C
            f["linear_bias"] += 2 * 2.0 * M * C * C + 2.0 * M * C * 8 * C + 2.0 * M * 4 * C * C   # 2 x to_out, FF
        else:
            single_layer(M, C, 4.0 * F * L * L * C)
        if T3d:
            f["conv1d_k1"] += 2 * 2.0 * M * C * C                         # proj_in_temporal, proj_out_temporal
            f["linear_nobias"] += 3 * 2.0 * M * C * C
            f["attention_temporal"] += 4.0 * (F // T3d) * L * T3d * T3d * C
            f["linear_bias"] += 2.0 * M * C * C + 2.0 * M * C * 8 * C + 2.0 * M * 4 * C * C
            if ca_type:
                lkv = {"center": L, "self": L, "center_self": 2 * L}[ca_type]
                f["conv1x1"] += 2 * 2.0 * M * C * C
                single_layer(M, C, 4.0 * F * L * lkv * C, Mkv=F * lkv)   # to_k/to_v run on the concatenated context

    def walk(blocks, F, Bemb, T3d, text=True, ca_type=None):
        for k, cin, cout, ds in blocks:
            L = (h // ds) * (w // ds)
            if k == "res":
                res(F, Bemb, L, cin, cout, T3d)
            elif k == "attn":
                attn(F, L, cin, T3d, text, ca_type)
            elif k == "down":
                f["conv3x3_s2"] += 2.0 * F * (L // 4) * 9 * cin * cout
                if T3d:
                    f["conv1d_k3"] += 2.0 * F * (L // 4) * 3 * cout * cout
            elif k == "up":
                f["conv3x3"] += 2.0 * F * 4 * L * 9 * cin * cout
                f["conv1d_k3"] += 2.0 * F * 4 * L * 3 * cout * cout

    def time_embed(Bn):
        f["linear_bias"] += 2.0 * Bn * (320 * emb_dim + emb_dim * emb_dim)

    def controlnet(F, Bn, hint_stem: bool, text: bool):
        """F frames through the 2-D encoder; Bn = batch entries seen by time_embed (before the per-frame repeat)."""
        time_embed(Bn)
        if hint_stem:                                                     # controlmodel.py:215-231 on the 8h x 8w hint
            L = 64 * h * w
            for i, s in enumerate(HINT_STRIDES):
                if s == 2:
                    L //= 4
                f["hint_stem"] += 2.0 * F * L * 9 * HINT_WIDTHS[i] * HINT_WIDTHS[i + 1]
        f["conv3x3"] += 2.0 * F * h * w * 9 * 4 * 320                     # input_blocks[0] (on x, or on cond_feat)
        walk(inp + mid, F, F, 0, text=text)
        f["conv1x1"] += 2.0 * F * h * w * 320 * 320                       # zero_convs[0]
        for k, _, c, ds in inp:                                           # one zero conv per input block output
            if k == "res":
                f["conv1x1"] += 2.0 * F * (h // ds) * (w // ds) * c * c
            elif k == "down":
                f["conv1x1"] += 2.0 * F * (h // (2 * ds)) * (w // (2 * ds)) * c * c
        f["conv1x1"] += 2.0 * F * (h // 8) * (w // 8) * 1280 * 1280       # middle_block_out

    F, L0 = B * T, h * w
    ca_type = "center_self" if kind == "tvi2v" else None
    # ControlNet2D on every frame (emb and context repeated per frame, controlmodel.py:260-266)
    controlnet(F, B, hint_stem=True, text=True)
    if kind == "tvi2v":                                                   # controlnet_img on the centre frame
        controlnet(B, B, hint_stem=False, text=False)
    # UNet3D
    time_embed(B)
    f["conv3x3"] += 2.0 * F * L0 * 9 * 4 * 320
    f["conv1d_k3"] += 2.0 * F * L0 * 3 * 320 * 320
    walk(inp + mid + out, F, B, T, text=True, ca_type=ca_type)
    f["conv3x3"] += 2.0 * F * L0 * 9 * 320 * 4
    f["conv1d_k3"] += 2.0 * F * L0 * 3 * 4 * 4
    f = dict(f)
    f["total"] = sum(f.values())
    return f


def decoder_flops(F: int, h: int, w: int, ch: int = 128, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2, z_channels: int = 4,
                  out_ch: int = 3) -> Dict[str, float]:
    """Algorithmic FLOPs (2 x MACs of every conv / bmm, as FlopCounterMode counts the reference module) of
    AutoencoderKL.decode on F frames of latent h x w: post_quant_conv + Decoder.forward (model.py:728-761)."""
    f = defaultdict(float)
    L = h * w
    block_in = ch * ch_mult[-1]
    f["conv1x1"] += 2.0 * F * L * z_channels * z_channels                                   # post_quant_conv
    f["conv3x3"] += 2.0 * F * L * 9 * z_channels * block_in                                 # conv_in

    def res(cin, cout, px):
        f["conv3x3"] += 2.0 * F * px * 9 * (cin * cout + cout * cout)
        if cin != cout:
            f["conv1x1"] += 2.0 * F * px * cin * cout

    res(block_in, block_in, L)
    f["conv1x1"] += 4 * 2.0 * F * L * block_in * block_in                                   # q, k, v, proj_out
    f["attention"] += 4.0 * F * L * L * block_in
    res(block_in, block_in, L)
    px = L
    for i_level in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[i_level]
        for _ in range(num_res_blocks + 1):
            res(block_in, block_out, px)
            block_in = block_out
        if i_level != 0:
            px *= 4
            f["conv3x3"] += 2.0 * F * px * 9 * block_in * block_in                          # Upsample.conv
    f["conv3x3"] += 2.0 * F * px * 9 * block_in * out_ch                                    # conv_out
    f = dict(f)
    f["total"] = sum(f.values())
    return f
