"""Frozen PSN backbone of DPMN's default configuration: TATT = `TSRN_TL_TRANS`.

The PSN is the caller-side step in front of the hot path (`interfaces/super_resolution.py:165-171`: `images_lr_psn, _ =
model_psn(images_lr, label_vecs)`); it is frozen (eval mode, no gradient), so what matters is its inference forward
(`model/tatt.py:645-691`).  This is an eval-only restatement on torch's own operators (cuDNN convolutions / GRU, the
fused multi-head attention) captured ONCE into a CUDA graph -- library kernels, deliberately: the PSN is outside the
PGRM / CMM hot path this package rebuilds, and the graph removes what actually costs time in the reference (some 400
eager launches per call).  Module names mirror the reference so that a `TSRN_TL_TRANS` checkpoint loads with
`load_reference_state_dict` (the STN head / TPS warp only run in training, `tatt.py:647-649`, and are not built).

Reference quirks kept (each pinned by tests/golden/tatt.npz, minted from the unmodified reference):
  * `InfoTransformer.gru_encoding` is `batch_first=True` but is fed (W, B, H*C): the recurrence runs over the BATCH axis
    (`transformer_v2.py:180,221`), so the query embedding of image b depends on b;
  * the encoder adds its input to itself before the first layer (`output + src_item` with `output = src`, `:276`);
  * decoder layers skip their self-attention (`TransformerDecoderLayer_TP.forward_post`, `:818-833`);
  * the text prior is the mean of the two decoder layers' normalised outputs (`tatt.py:218`).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _mish(x: torch.Tensor) -> torch.Tensor:                      # tatt.py:1055-1063
    return x * torch.tanh(F.softplus(x))


class GruBlock(nn.Module):
    """1x1 conv, then a bidirectional GRU along the LAST spatial axis, one sequence per row (tatt.py:1066-1083)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.gru = nn.GRU(out_channels, out_channels // 2, bidirectional=True, batch_first=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.conv1(x).permute(0, 2, 3, 1)                    # (B, H, W, C)
        B, H, W, C = y.shape
        y, _ = self.gru(y.reshape(B * H, W, C))
        return y.reshape(B, H, W, C).permute(0, 3, 1, 2)


class RecurrentResidualBlockTL(nn.Module):
    """SRB with the text prior concatenated in front of the first (vertical) GRU (tatt.py:873-909)."""

    def __init__(self, channels: int, text_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.gru1 = GruBlock(channels + text_channels, channels)
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)
        self.gru2 = GruBlock(channels, channels)

    def forward(self, x: torch.Tensor, text_emb: torch.Tensor) -> torch.Tensor:
        r = self.bn2(self.conv2(_mish(self.bn1(self.conv1(x)))))
        r = self.gru1(torch.cat([r, text_emb], 1).transpose(-1, -2)).transpose(-1, -2)
        return self.gru2(x + r)


class _EncoderLayer(nn.Module):                                  # transformer_v2.py:448-484 (post-norm)
    def __init__(self, d: int, heads: int, ff: int):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, heads)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)

    def forward(self, src: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        qk = src + pos
        src = self.norm1(src + self.self_attn(qk, qk, src, need_weights=False)[0])
        return self.norm2(src + self.linear2(F.relu(self.linear1(src))))


class _DecoderLayerTP(nn.Module):                                # transformer_v2.py:773-833 (post-norm, no self-attention)
    def __init__(self, d: int, heads: int, ff: int):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, heads)         # parameters exist in the reference; never called
        self.multihead_attn = nn.MultiheadAttention(d, heads)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)                             # unused, as in the reference
        self.norm2 = nn.LayerNorm(d)
        self.norm3 = nn.LayerNorm(d)

    def forward(self, tgt, memory, pos, query_pos):
        t2, w = self.multihead_attn(tgt + query_pos, memory + pos, memory)
        tgt = self.norm2(tgt + t2)
        return self.norm3(tgt + self.linear2(F.relu(self.linear1(tgt)))), w


class _Stack(nn.Module):
    def __init__(self, layers, norm: Optional[nn.Module] = None):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        if norm is not None:
            self.norm = norm


class InfoTransformer(nn.Module):                                # transformer_v2.py:154-245
    def __init__(self, d: int, heads: int, ff: int, n_enc: int, n_dec: int, feat_h: int, feat_w: int):
        super().__init__()
        self.encoder = _Stack([_EncoderLayer(d, heads, ff) for _ in range(n_enc)])
        self.decoder = _Stack([_DecoderLayerTP(d, heads, ff) for _ in range(n_dec)], nn.LayerNorm(d))
        self.gru_encoding = nn.GRU(d * feat_h, d * feat_h // 2, bidirectional=True, batch_first=True)
        self.feat = (feat_h, feat_w)

    def forward(self, src, query_embed, pos_embed, tgt):
        fh, fw = self.feat
        _, bs, d = src.shape
        q = query_embed.unsqueeze(1).expand(-1, bs, -1).reshape(fh, fw, bs, d).permute(1, 2, 0, 3).reshape(fw, bs, fh * d)
        q, _ = self.gru_encoding(q.contiguous())                 # batch_first GRU on (W, B, H*C): recurrence over B
        q = q.reshape(fw, bs, fh, d).permute(2, 0, 1, 3).reshape(fh * fw, bs, d)
        memory = src
        for layer in self.encoder.layers:
            memory = layer(memory + src, pos_embed)              # `output + src_item` (transformer_v2.py:276)
        outs, w = [], None
        out = tgt
        for layer in self.decoder.layers:
            out, w = layer(out, memory, pos_embed, q)
            outs.append(self.decoder.norm(out))
        return torch.stack(outs), w


class _PE(nn.Module):                                            # transformer_v2.py:22-42 (dropout is identity in eval)
    def __init__(self, d: int, max_len: int = 5000):
        super().__init__()
        pe = torch.zeros(max_len, d)
        position = torch.arange(0, max_len).unsqueeze(1).float()
        div = torch.exp(torch.arange(0, d, 2).float() * -(math.log(10000.0) / d))
        pe[:, 0::2] = torch.sin(position * div)
        pe[:, 1::2] = torch.cos(position * div)
        self.register_buffer("pe", pe.unsqueeze(0))


class TPInterpreter(nn.Module):                                  # tatt.py:154-223
    def __init__(self, t_emb: int, d: int, output_size=(16, 64), feature_in: int = 64):
        super().__init__()
        self.fc_in = nn.Linear(t_emb, d)
        self.fc_feature_in = nn.Linear(feature_in, d)            # in the reference's state_dict, never called
        self.activation = nn.PReLU()
        self.upsample_transformer = InfoTransformer(d, 4, d, 1, 2, output_size[0], output_size[1])
        self.pe = _PE(d)
        self.init_factor = nn.Embedding(output_size[0] * output_size[1], d)

    def forward(self, image_feature: torch.Tensor, tp_input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        N, C, H, W = image_feature.shape
        x_im = image_feature.reshape(N, C, H * W).permute(2, 0, 1)
        x = self.activation(self.fc_in(tp_input.permute(0, 3, 1, 2).squeeze(-1)))    # (N, 26, d)
        Lt = x.shape[1]
        pos = self.pe.pe[0, :Lt].unsqueeze(1).expand(-1, N, -1)                      # pe(zeros): the table itself
        hs, w = self.upsample_transformer(x.permute(1, 0, 2), self.init_factor.weight, pos, x_im)
        return hs.mean(0).permute(1, 2, 0).reshape(N, C, H, W), w


class UpsampleBLock(nn.Module):                                  # tatt.py:1039-1052
    def __init__(self, ch: int, up: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch * up * up, kernel_size=3, padding=1)
        self.up = up

    def forward(self, x):
        return _mish(F.pixel_shuffle(self.conv(x), self.up))


class TATT(nn.Module):
    """`TSRN_TL_TRANS(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)` in eval
    mode (interfaces/base.py:145-148): (B, 4, 16, 64) LR image + (B, 37, 1, 26) text prior -> (B, 4, 32, 128), attention
    weights (B, 1024, 26)."""

    def __init__(self, scale_factor: int = 2, width: int = 128, height: int = 32, srb_nums: int = 5, mask: bool = True,
                 hidden_units: int = 32, text_emb: int = 37, out_text_channels: int = 64):
        super().__init__()
        in_planes = 4 if mask else 3
        ch = 2 * hidden_units
        self.srb_nums = srb_nums
        self.block1 = nn.Sequential(nn.Conv2d(in_planes, ch, kernel_size=9, padding=4), nn.PReLU())
        for i in range(srb_nums):
            setattr(self, f"block{i + 2}", RecurrentResidualBlockTL(ch, out_text_channels))
        self.infoGen = TPInterpreter(text_emb, out_text_channels, (height // scale_factor, width // scale_factor))
        setattr(self, f"block{srb_nums + 2}", nn.Sequential(nn.Conv2d(ch, ch, kernel_size=3, padding=1), nn.BatchNorm2d(ch)))
        ups = [UpsampleBLock(ch, 2) for _ in range(int(math.log2(scale_factor)))]
        setattr(self, f"block{srb_nums + 3}", nn.Sequential(*ups, nn.Conv2d(ch, in_planes, kernel_size=9, padding=4)))
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)
        self._graph = None

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("dpmn_b200.psn.TATT is the FROZEN PSN: eval-only (the STN / TPS training path is not built)")
        return super().train(False)

    def load_reference_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Load a reference `TSRN_TL_TRANS` state_dict; its training-only STN / TPS entries are skipped."""
        keep = {k: v for k, v in sd.items() if not (k.startswith("stn_head.") or k.startswith("tps."))}
        own = self.state_dict()
        # the reference's UpsampleBLock has no parameters besides `conv`; names already coincide
        missing = [k for k in own if k not in keep]
        extra = [k for k in keep if k not in own]
        if missing or extra:
            raise KeyError(f"TATT state_dict mismatch: missing {missing[:5]}, unexpected {extra[:5]}")
        self.load_state_dict(keep, strict=True)
        self._graph = None          # captured graphs bake the old weights' addresses only; values are read at replay -- reset anyway

    def forward(self, x: torch.Tensor, text_emb: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if text_emb is None:
            text_emb = torch.zeros(1, 37, 1, 26, device=x.device, dtype=x.dtype)     # tatt.py:651-652
        b1 = self.block1(x)
        tp_map, pr_weights = self.infoGen(b1, text_emb)
        f = b1
        for i in range(self.srb_nums):
            f = getattr(self, f"block{i + 2}")(f, tp_map)
        f = getattr(self, f"block{self.srb_nums + 2}")(f)
        return torch.tanh(getattr(self, f"block{self.srb_nums + 3}")(b1 + f)), pr_weights

    # ---- CUDA-graph replay (static batch) ------------------------------------------------------------------------------
    @torch.no_grad()
    def graphed(self, x: torch.Tensor, text_emb: torch.Tensor, slot: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Same forward, replayed on the current stream from a CUDA graph captured on first use for this (batch, device, slot);
        inputs are copied into the slot's static buffers, the returned tensors are its static outputs (valid until the
        slot's next call).  Different slots own different buffers, so they may be in flight on different streams."""
        if not x.is_cuda:
            raise RuntimeError("TATT.graphed needs CUDA tensors")
        key = (tuple(x.shape), tuple(text_emb.shape), x.device)
        if self._graph is None:
            self._graph = {}
        ent = self._graph.get(slot)
        if ent is None or ent[0] != key:
            sx, st = x.clone(), text_emb.clone()
            cur = torch.cuda.current_stream(x.device)
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):                               # warm-up: cuDNN plans, GRU weight flattening
                    self.forward(sx, st)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    out = self.forward(sx, st)
            cur.wait_stream(side)
            ent = (key, g, sx, st, out)
            self._graph[slot] = ent
        _, g, sx, st, out = ent
        sx.copy_(x, non_blocking=True)
        st.copy_(text_emb, non_blocking=True)
        g.replay()
        return out
