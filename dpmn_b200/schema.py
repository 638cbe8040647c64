"""state_dict schemas of the two hot-path modules (the weight contract of the drop-in boundary).

The reference has no FFI layer; what a checkpoint written by the reference must find on our side is
the exact set of parameter / buffer names and shapes.  These functions enumerate them.

  PGRM  : /root/reference/model/pgrm.py:462-522 (ctor), :109-182 (WindowAttention buffers)
  CMM   : /root/reference/model/cmm.py:80-118

Every entry is (name, shape, kind) with kind in {"param", "buffer"}.  `tests/test_schema.py` checks
these lists against the committed golden dump of the reference's own `state_dict()` keys.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

Entry = Tuple[str, Tuple[int, ...], str]


@dataclass(frozen=True)
class PGRMConfig:
    """Resolved (per-`iter`) hyper-parameters of one PGRM (pgrm.py:462-522)."""
    img_size: Tuple[int, int] = (32, 128)
    patch_size: int = 2
    in_chans: int = 3
    embed_dim: int = 96
    num_heads: int = 6
    window_size: Tuple[int, ...] = (2, 4, 8)
    mlp_ratio: float = 4.0
    iter: int = 0
    mode: bool = True            # False -> has prior_fusion conv (2-channel rendered-text prior)
    hidden_size: int = 3
    depth: int = 2               # BasicLayer depth is hard-wired to 2 (pgrm.py:506)

    @property
    def grid(self) -> Tuple[int, int]:
        return (self.img_size[0] // self.patch_size, self.img_size[1] // self.patch_size)

    @property
    def tokens(self) -> int:
        return self.grid[0] * self.grid[1]

    @property
    def groups(self) -> int:
        return len(self.window_size)

    @property
    def group_channels(self) -> int:
        return self.embed_dim // self.groups

    @property
    def heads_per_group(self) -> int:
        return self.num_heads // self.groups

    @property
    def head_dim(self) -> int:
        return self.group_channels // self.heads_per_group

    @property
    def mlp_hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)

    def effective_windows(self, block: int) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
        """(window sizes, shift sizes) of block `block` after the clamp at pgrm.py:147-151.

        block 0 has shift 0, block 1 has shift ws//2 (pgrm.py:362); a window that is >= min(H, W)
        is clamped to min(H, W) with shift 0.
        """
        H, W = self.grid
        ws_out, sh_out = [], []
        for ws in self.window_size:
            sh = 0 if block % 2 == 0 else ws // 2
            if min(H, W) <= ws:
                sh = 0
                ws_eff = min(H, W)
            else:
                ws_eff = ws
            ws_out.append(ws_eff)
            sh_out.append(sh)
        return tuple(ws_out), tuple(sh_out)


def pgrm_schema(cfg: PGRMConfig) -> List[Entry]:
    C = cfg.embed_dim
    G = cfg.groups
    hg = cfg.heads_per_group
    cg = cfg.group_channels
    hid = cfg.mlp_hidden
    p = cfg.patch_size
    H, W = cfg.grid
    out: List[Entry] = []
    for i in range(cfg.iter + 1):
        out.append((f"weight_list_{i}", (1, cfg.hidden_size, 32, 128), "param"))  # pgrm.py:496-497
    if not cfg.mode:
        out.append(("prior_fusion.weight", (3, 2, 3, 3), "param"))
        out.append(("prior_fusion.bias", (3,), "param"))
    out.append(("patch_embed.proj.weight", (C, cfg.in_chans, p, p), "param"))
    out.append(("patch_embed.proj.bias", (C,), "param"))
    out.append(("patch_embed.norm.weight", (C,), "param"))
    out.append(("patch_embed.norm.bias", (C,), "param"))
    for b in range(cfg.depth):
        pre = f"layers.0.blocks.{b}."
        ws_eff, sh_eff = cfg.effective_windows(b)
        out.append((pre + "norm1_q.weight", (C,), "param"))
        out.append((pre + "norm1_q.bias", (C,), "param"))
        out.append((pre + "norm1_kv.weight", (C,), "param"))
        out.append((pre + "norm1_kv.bias", (C,), "param"))
        for g, ws in enumerate(cfg.window_size):
            # tables and index buffers are sized by the UN-clamped window (pgrm.py:127-145)
            out.append((pre + f"attn.relative_position_bias_table_{g}", ((2 * ws - 1) ** 2, hg), "param"))
        for g, ws in enumerate(cfg.window_size):
            out.append((pre + f"attn.relative_position_index_{g}", (ws * ws, ws * ws), "buffer"))
        for g in range(G):
            if sh_eff[g] > 0:
                nW = (H // ws_eff[g]) * (W // ws_eff[g])
                n = ws_eff[g] * ws_eff[g]
                out.append((pre + f"attn.attn_mask_{g}", (nW, n, n), "buffer"))
        out.append((pre + "attn.q.weight", (C, C), "param"))
        out.append((pre + "attn.q.bias", (C,), "param"))
        out.append((pre + "attn.kv.weight", (2 * C, C), "param"))
        out.append((pre + "attn.kv.bias", (2 * C,), "param"))
        out.append((pre + "attn.sknet.proj.weight", (C, C), "param"))
        out.append((pre + "attn.sknet.proj.bias", (C,), "param"))
        out.append((pre + "attn.sknet.fc1.weight", (cg // 2, C), "param"))
        out.append((pre + "attn.sknet.fc1.bias", (cg // 2,), "param"))
        out.append((pre + "attn.sknet.fc2.weight", (C, cg // 2), "param"))
        out.append((pre + "attn.sknet.fc2.bias", (C,), "param"))
        out.append((pre + "attn.sknet.proj_head.weight", (C, cg), "param"))
        out.append((pre + "attn.sknet.proj_head.bias", (C,), "param"))
        out.append((pre + "norm2.weight", (C,), "param"))
        out.append((pre + "norm2.bias", (C,), "param"))
        out.append((pre + "mlp.fc1.weight", (hid, C), "param"))
        out.append((pre + "mlp.fc1.bias", (hid,), "param"))
        out.append((pre + "mlp.fc2.weight", (C, hid), "param"))
        out.append((pre + "mlp.fc2.bias", (C,), "param"))
        out.append((pre + "mlp.depthwise_conv.weight", (hid, 1, 3, 3), "param"))
        out.append((pre + "mlp.depthwise_conv.bias", (hid,), "param"))
        out.append((pre + "mlp.pointwise_conv.weight", (hid, hid, 1, 1), "param"))
        out.append((pre + "mlp.pointwise_conv.bias", (hid,), "param"))
    hp = cfg.hidden_size * p * p
    out.append(("conv_before_upsample.0.weight", (hp, C, 3, 3), "param"))
    out.append(("conv_before_upsample.0.bias", (hp,), "param"))
    out.append(("conv_before_upsample.1.weight", (hp, hp, 3, 3), "param"))
    out.append(("conv_before_upsample.1.bias", (hp,), "param"))
    return out


def _bn(prefix: str, ch: int) -> List[Entry]:
    return [
        (prefix + ".weight", (ch,), "param"),
        (prefix + ".bias", (ch,), "param"),
        (prefix + ".running_mean", (ch,), "buffer"),
        (prefix + ".running_var", (ch,), "buffer"),
        (prefix + ".num_batches_tracked", (), "buffer"),
    ]


def cmm_schema(c_img: int = 3, cnum: int = 64) -> List[Entry]:
    """cmm.py:80-118.  EncodeBlock = [act, conv4x4s2d2p3 (in->in), BN, act, conv3x3 (in->out), BN]
    (indices 1,2,4,5 hold state); DecodeBlock = [act, convT3x3 (in->out), BN, act, convT4x4s2 (out->out), BN]."""
    out: List[Entry] = []

    def conv(name, o, i, k):
        out.append((name + ".weight", (o, i, k, k), "param"))
        out.append((name + ".bias", (o,), "param"))

    def convT(name, i, o, k):
        out.append((name + ".weight", (i, o, k, k), "param"))
        out.append((name + ".bias", (o,), "param"))

    enc_ch = [(cnum, cnum * 2), (cnum * 2, cnum * 4), (cnum * 4, cnum * 8), (cnum * 8, cnum * 8)]
    for br in (1, 2):
        conv(f"en_1_{br}", cnum, c_img, 3)
        for lvl, (ci, co) in zip((2, 3, 4, 5), enc_ch):
            pre = f"en_{lvl}_{br}.encode."
            conv(pre + "1", ci, ci, 4)
            out.extend(_bn(pre + "2", ci))
            conv(pre + "4", co, ci, 3)
            out.extend(_bn(pre + "5", co))
        conv(f"en_6_{br}.1", cnum * 8, cnum * 8, 4)
    out.append(("fc_1.weight", (4 * cnum, 16 * cnum), "param"))
    out.append(("fc_1.bias", (4 * cnum,), "param"))
    out.append(("fc_2.weight", (16 * cnum, 4 * cnum), "param"))
    out.append(("fc_2.bias", (16 * cnum,), "param"))
    convT("de_6.1", cnum * 16, cnum * 8, 4)
    out.extend(_bn("de_6.2", cnum * 8))
    dec_ch = [(5, cnum * 8 * 3, cnum * 8), (4, cnum * 8 * 3, cnum * 4), (3, cnum * 4 * 3, cnum * 2), (2, cnum * 2 * 3, cnum)]
    for lvl, ci, co in dec_ch:
        pre = f"de_{lvl}.decode."
        convT(pre + "1", ci, co, 3)
        out.extend(_bn(pre + "2", co))
        convT(pre + "4", co, co, 4)
        out.extend(_bn(pre + "5", co))
    convT("de_1.1", cnum * 3, c_img, 3)
    return out


def resolve_pgrm_config(img_size=(32, 128), patch_size=(2,), in_chans=3, embed_dim=(96,), depths=(1,),
                        num_heads=((6,),), window_size=((2, 4, 8),), mlp_ratio=(4.,), iter=0, mode=True,
                        hidden_size=64) -> PGRMConfig:
    """Index the reference's list-valued ctor args by `iter` (pgrm.py:472-477,507-508)."""
    if depths[iter] != 1:
        # depths[iter] > 1 builds a second BasicLayer of width 2*embed_dim with downsample=None
        # (pgrm.py:503-514), which cannot consume the first layer's output in the reference either.
        raise ValueError("PGRM: depths[iter] must be 1 (the reference cannot run deeper stacks)")
    return PGRMConfig(img_size=(int(img_size[0]), int(img_size[1])), patch_size=int(patch_size[iter]),
                      in_chans=int(in_chans), embed_dim=int(embed_dim[iter]),
                      num_heads=int(num_heads[iter][0]), window_size=tuple(int(w) for w in window_size[iter]),
                      mlp_ratio=float(mlp_ratio[iter]), iter=int(iter), mode=bool(mode),
                      hidden_size=int(hidden_size))
