"""Drop-in mirrors of the reference's model classes (same class names, constructor arguments, forward signatures and
``state_dict`` keys) whose forwards run on the sm_100a library through the C ABI.

    reference                                  here
    model/model.py:16-359     MESM             MESM               -> mesm_forward
    model/transformer.py:62   T2VEncoder       T2VEncoder         -> mesm_t2v_encoder
    model/transformer.py:108  T2VEncoder_TwoMLP T2VEncoder_TwoMLP
    model/transformer.py:119  Transformer      Transformer        -> mesm_transformer
    model/attention.py:61     MultiheadAttention MultiheadAttention -> mesm_mha_noproj
    model/position_encoding.py:35 PositionEmbeddingSine (parameter-free; the fused forward computes it on the fly)
    model/text_encoder.py:240 CLIPTextEncoder  CLIPTextEncoder    -> mesm_clip_forward
    runner.py:255-298         build_model      build_model

The nn.Modules below only *hold parameters* under the reference's names; no torch op touches the hot path.  Inference
(eval mode) only: dropout is the identity and the training-only MLM branch (model/model.py:307-332) is not provided.
"""
import ctypes
import math
from ctypes import byref
from typing import Optional

import torch
from torch import nn

from . import _lib
from .engine import Engine, _f32, _ptr, _stream, _u8
from ._lib import check


# ---------------------------------------------------------------------------------------------------------------------
# parameter holders (names = reference attribute names)
# ---------------------------------------------------------------------------------------------------------------------
class MLP(nn.Module):
    """model/model.py:397-409 / model/transformer.py:21-33 (parameters only)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


class LinearLayer(nn.Module):
    """model/model.py:412-434 (parameters only): LayerNorm -> Dropout -> Linear -> optional ReLU."""

    def __init__(self, in_hsz, out_hsz, layer_norm=True, dropout=0.1, relu=True):
        super().__init__()
        self.relu = relu
        self.layer_norm = layer_norm
        if layer_norm:
            self.LayerNorm = nn.LayerNorm(in_hsz)
        self.net = nn.Sequential(nn.Dropout(dropout), nn.Linear(in_hsz, out_hsz))


class _PackedMHA(nn.Module):
    """Parameter layout of nn.MultiheadAttention(d, h): in_proj_weight [3d,d], in_proj_bias [3d], out_proj."""

    def __init__(self, d_model):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


def _activation(name):
    if name != "prelu":
        raise NotImplementedError("mesm_b200 implements activation='prelu' (what runner.py:199,209,221,235 passes)")
    return nn.PReLU()


class T2V_TransformerEncoderLayer(nn.Module):
    """model/transformer.py:485-503 (parameters only)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before:
            raise NotImplementedError("self.normalize_before is True")      # same as transformer.py:552-553
        self.self_attn = _PackedMHA(d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.activation = _activation(activation)
        self.nhead = nhead


class T2V_TransformerEncoderLayer_TwoMLP(T2V_TransformerEncoderLayer):
    """model/transformer.py:562-571: second FFN / norms used only when is_MLM=True (training)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.linear1_1 = nn.Linear(d_model, dim_feedforward)
        self.linear2_1 = nn.Linear(dim_feedforward, d_model)
        self.norm1_1 = nn.LayerNorm(d_model)
        self.norm2_1 = nn.LayerNorm(d_model)


class _LayerStack(nn.Module):
    def __init__(self, make_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList(make_layer() for _ in range(num_layers))
        self.num_layers = num_layers
        self.norm = None


TransformerEncoderLayer = T2V_TransformerEncoderLayer      # same parameter set (transformer.py:615-632)


class TransformerDecoderLayer(nn.Module):
    """model/transformer.py:676-718 (parameters only)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", keep_query_pos=False):
        super().__init__()
        for n in ("sa_qcontent_proj", "sa_qpos_proj", "sa_kcontent_proj", "sa_kpos_proj", "sa_v_proj",
                  "ca_qcontent_proj", "ca_qpos_proj", "ca_kcontent_proj", "ca_kpos_proj", "ca_v_proj", "ca_qpos_sine_proj"):
            setattr(self, n, nn.Linear(d_model, d_model))
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout, vdim=d_model)
        self.cross_attn = MultiheadAttention(d_model * 2, nhead, dropout=dropout, vdim=d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.activation = _activation(activation)


class _Decoder(nn.Module):
    """model/transformer.py:280-331 (parameters only)."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout, activation, num_layers):
        super().__init__()
        self.layers = nn.ModuleList(TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation)
                                    for _ in range(num_layers))
        self.num_layers = num_layers
        self.norm = nn.LayerNorm(d_model)
        self.query_scale = MLP(d_model, d_model, d_model, 2)
        self.ref_point_head = MLP(d_model, d_model, d_model, 2)
        self.bbox_embed = MLP(d_model, d_model, 2, 3)
        nn.init.constant_(self.bbox_embed.layers[-1].weight.data, 0)
        nn.init.constant_(self.bbox_embed.layers[-1].bias.data, 0)
        self.ref_anchor_head = MLP(d_model, d_model, 1, 2)
        for layer_id in range(num_layers - 1):
            self.layers[layer_id + 1].ca_qpos_proj = None            # transformer.py:329-331


def _xavier(module):
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)


# ---------------------------------------------------------------------------------------------------------------------
# engine-backed modules
# ---------------------------------------------------------------------------------------------------------------------
class _EngineBacked(nn.Module):
    """Owns an Engine whose weights mirror this module's parameters under ``_prefix``; re-synced when they change."""
    _prefix = ""

    def _engine_cfg(self):
        raise NotImplementedError

    def _weights_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _engine(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("mesm_b200: inputs must be CUDA tensors (there is no CPU fallback)")
        sig = (str(device), self._weights_signature())
        eng = getattr(self, "_eng", None)
        if eng is None or self.__dict__.get("_eng_sig") != sig:
            if eng is None or str(eng.device) != str(device):
                eng = Engine(self._engine_cfg(), device=device, chunk_pairs=self.__dict__.get("chunk_pairs", 0))
            # (checkpoints omit text_encoder.*, utils/model_utils.py:20-36; a CLIP / GloVe front-end keeps its own weights)
            eng.load_state_dict({self._prefix + k: v for k, v in self.state_dict().items() if not k.startswith("text_encoder.")})
            self.__dict__["_eng"] = eng
            self.__dict__["_eng_sig"] = (str(device), self._weights_signature())
        return eng


class T2VEncoder(_EngineBacked):
    """model/transformer.py:62-105.  forward(src_txt[B,Lt,d], src_vid[B,Lv,d], src_txt_mask, src_txt_key_padding_mask
    [B,Lt] (True = pad), pos_txt, src_vid_mask, src_vid_key_padding_mask [B,Lv], pos_vid) -> [B,Lv,d]."""
    _prefix = "t2v_encoder."
    _layer_cls = T2V_TransformerEncoderLayer

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, dim_feedforward=2048, dropout=0.1, activation="relu",
                 normalize_before=False):
        super().__init__()
        self.t2v_encoder = _LayerStack(
            lambda: self._layer_cls(d_model, nhead, dim_feedforward, dropout, activation, normalize_before),
            num_encoder_layers)
        _xavier(self)
        self.d_model, self.nhead, self.dim_feedforward = d_model, nhead, dim_feedforward

    def _engine_cfg(self):
        return dict(v_feat_dim=4, t_feat_dim=4, hidden_dim=self.d_model, nheads=self.nhead,
                    dim_feedforward=self.dim_feedforward, t2v_layers=self.t2v_encoder.num_layers, num_recfw_layers=0)

    def forward(self, src_txt, src_vid, src_txt_mask=None, src_txt_key_padding_mask=None, pos_txt=None,
                src_vid_mask=None, src_vid_key_padding_mask=None, pos_vid=None, **kwargs):
        if kwargs.get("is_MLM"):
            raise NotImplementedError("is_MLM=True is the training-only MLM branch (model/model.py:327-331)")
        eng = self._engine(src_vid.device)
        B, Lv, _ = src_vid.shape
        Lt = src_txt.shape[1]
        txt, vid = _f32(src_txt, "src_txt"), _f32(src_vid, "src_vid")
        tp, vp = _u8(src_txt_key_padding_mask, "src_txt_key_padding_mask"), _u8(src_vid_key_padding_mask, "src_vid_key_padding_mask")
        pt = None if pos_txt is None else _f32(pos_txt, "pos_txt")
        pv = None if pos_vid is None else _f32(pos_vid, "pos_vid")
        out = torch.empty(B, Lv, self.d_model, dtype=torch.float32, device=vid.device)
        lib = eng.lib
        with torch.cuda.device(vid.device):
            ws = eng._workspace(lib.mesm_t2v_workspace_bytes(B, Lt, Lv))
            check(lib.mesm_t2v_encoder(eng.ctx, b"t2v_encoder", _ptr(txt), _ptr(vid), _ptr(tp), _ptr(vp), _ptr(pt), _ptr(pv),
                                       B, Lt, Lv, _ptr(out), _ptr(ws), ws.numel(), _stream()), eng.ctx)
        return out


class T2VEncoder_TwoMLP(T2VEncoder):
    """model/transformer.py:108-116."""
    _layer_cls = T2V_TransformerEncoderLayer_TwoMLP


class MultiheadAttention(_EngineBacked):
    """model/attention.py:61-182: projection-free MHA (q/k/v already projected by the caller), out_proj(vdim, vdim).
    forward(query[L,B,E], key[S,B,E], value[S,B,vdim], key_padding_mask[B,S], need_weights, attn_mask) ->
    (out[L,B,vdim], head-averaged weights [B,L,S] or None)."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None,
                 vdim=None):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.vdim = vdim if vdim is not None else embed_dim
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        if add_bias_kv or add_zero_attn:
            raise NotImplementedError("add_bias_kv / add_zero_attn are never used by MESM")
        self.out_proj = nn.Linear(self.vdim, self.vdim)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None):
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is always None at the reference's call sites (transformer.py:749,786)")
        lib = _lib.lib()
        q, k, v = _f32(query, "query"), _f32(key, "key"), _f32(value, "value")
        L, B, E = q.shape
        S = k.shape[0]
        kp = None if key_padding_mask is None else _u8(key_padding_mask, "key_padding_mask")
        out = torch.empty(L, B, self.vdim, dtype=torch.float32, device=q.device)
        w = torch.empty(B, L, S, dtype=torch.float32, device=q.device) if need_weights else None
        ow, ob = _f32(self.out_proj.weight.detach(), "out_proj.weight"), _f32(self.out_proj.bias.detach(), "out_proj.bias")
        ws = torch.empty(lib.mesm_mha_workspace_bytes(L, S, B, E, self.vdim), dtype=torch.uint8, device=q.device)
        with torch.cuda.device(q.device):
            check(lib.mesm_mha_noproj(_ptr(q), _ptr(k), _ptr(v), L, S, B, E, self.vdim, self.num_heads, _ptr(ow), _ptr(ob),
                                      _ptr(kp), _ptr(out), _ptr(w), _ptr(ws), ws.numel(), _stream()))
        return out, w


class Transformer(_EngineBacked):
    """model/transformer.py:119-205.  forward(src[B,L,d], mask[B,L] (True = pad), query_embed[nq,2], pos_embed[B,L,d],
    global_token[B,1,d], global_token_pos[B,1,d]) -> (hs[nl,B,nq,d], references[nl,B,nq,2], memory_local[B,L,d],
    memory_global[B,d])."""
    _prefix = "transformer."

    def __init__(self, d_model=512, nhead=8, num_queries=2, num_encoder_layers=6, num_decoder_layers=6,
                 dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False,
                 return_intermediate_dec=False, query_dim=2, keep_query_pos=False, query_scale_type='cond_elewise',
                 num_patterns=0, modulate_t_attn=True, bbox_embed_diff_each_layer=False):
        super().__init__()
        if normalize_before or keep_query_pos or query_scale_type != 'cond_elewise' or not modulate_t_attn \
                or bbox_embed_diff_each_layer or query_dim != 2:
            raise NotImplementedError("only the configuration runner.py:225-236 builds is implemented")
        self.encoder = _LayerStack(lambda: TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation),
                                   num_encoder_layers)
        self.decoder = _Decoder(d_model, nhead, dim_feedforward, dropout, activation, num_decoder_layers)
        _xavier(self)
        self.d_model, self.nhead, self.dec_layers, self.num_queries = d_model, nhead, num_decoder_layers, num_queries
        self.dim_feedforward, self.dropout, self.activation, self.normalize_before = dim_feedforward, dropout, activation, normalize_before
        self.num_decoder_layers, self.num_encoder_layers = num_decoder_layers, num_encoder_layers

    def _engine_cfg(self):
        return dict(v_feat_dim=4, t_feat_dim=4, hidden_dim=self.d_model, nheads=self.nhead, t2v_layers=0, num_recfw_layers=0,
                    dim_feedforward=self.dim_feedforward, enc_layers=self.num_encoder_layers, dec_layers=self.num_decoder_layers,
                    num_queries=self.__dict__.get("_nq", 10))

    def forward(self, src, mask, query_embed, pos_embed, global_token, global_token_pos):
        self.__dict__["_nq"] = int(query_embed.shape[0])
        eng = self._engine(src.device)
        B, L, d = src.shape
        nq, nl = query_embed.shape[0], self.num_decoder_layers
        s, pe = _f32(src, "src"), _f32(pos_embed, "pos_embed")
        qe = _f32(query_embed, "query_embed")
        gt = _f32(global_token.reshape(-1, d)[0], "global_token")
        gp = _f32(global_token_pos.reshape(-1, d)[0], "global_token_pos")
        pad = _u8(mask, "mask")
        f = lambda *sh: torch.empty(*sh, dtype=torch.float32, device=s.device)
        hs, refs, mem, memg = f(nl, B, nq, d), f(nl, B, nq, 2), f(B, L, d), f(B, d)
        lib = eng.lib
        with torch.cuda.device(s.device):
            ws = eng._workspace(lib.mesm_transformer_workspace_bytes(eng.ctx, B, L))
            check(lib.mesm_transformer(eng.ctx, _ptr(s), _ptr(pad), _ptr(qe), _ptr(pe), _ptr(gt), _ptr(gp), B, L, _ptr(hs),
                                       _ptr(refs), _ptr(mem), _ptr(memg), _ptr(ws), ws.numel(), _stream()), eng.ctx)
        return hs, refs, mem, memg


class PositionEmbeddingSine(nn.Module):
    """model/position_encoding.py:35-72 (parameter-free).  MESM.forward computes it inside the library; this standalone
    forward exists for callers of the sub-modules and is plain tensor plumbing."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats, self.temperature, self.normalize = num_pos_feats, temperature, normalize
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, x, mask):
        x_embed = mask.cumsum(1, dtype=torch.float32)
        if self.normalize:
            x_embed = x_embed / (x_embed[:, -1:] + 1e-6) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=x.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode='trunc') / self.num_pos_feats)
        pos_x = x_embed[:, :, None] / dim_t
        return torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).flatten(2)


class TrainablePositionalEncoding(nn.Module):
    """model/position_encoding.py:10-32 (parameters only; use_txt_pos is false in every shipped config)."""

    def __init__(self, max_position_embeddings, hidden_size, dropout=0.1):
        super().__init__()
        self.position_embeddings = nn.Embedding(max_position_embeddings, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size)


class SegSenRecon(nn.Module):
    """model/model.py:437-465 (parameters only)."""

    def __init__(self, input_dropout, hidden_dim=512, nhead=8, num_layers=6, dim_feedforward=2048, dropout=0.1,
                 activation="relu", normalize_before=False):
        super().__init__()
        self.masked_sent_token = nn.Parameter(torch.zeros(hidden_dim).float(), requires_grad=True)
        self.recon_trans = _LayerStack(
            lambda: T2V_TransformerEncoderLayer(hidden_dim, nhead, dim_feedforward, dropout, activation, normalize_before),
            num_layers)
        self.output_sent_proj = nn.Sequential(
            LinearLayer(hidden_dim, hidden_dim, layer_norm=True, dropout=input_dropout, relu=True),
            LinearLayer(hidden_dim, hidden_dim, layer_norm=True, dropout=input_dropout, relu=False))


class GloveTextEncoder(nn.Module):
    """model/text_encoder.py:432-454: frozen embedding lookup (the gather is plumbing in front of the path)."""

    def __init__(self, vocab_size, embed_dim):
        super().__init__()
        self.emb = nn.Embedding(vocab_size, embed_dim)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, word_ids):
        return self.emb(word_ids)


class _ClipBlock(nn.Module):
    """ResidualAttentionBlock (model/text_encoder.py:165-186), parameters only."""

    def __init__(self, d_model):
        super().__init__()
        self.attn = _PackedMHA(d_model)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(d_model, d_model * 4))
        self.mlp.add_module("c_proj", nn.Linear(d_model * 4, d_model))
        self.ln_2 = nn.LayerNorm(d_model)


class _ClipTransformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.resblocks = nn.Sequential(*[_ClipBlock(width) for _ in range(layers)])


class CLIPTextEncoder(nn.Module):
    """Drop-in for model/text_encoder.py:240-354 (same constructor, state_dict keys and output dict).  forward(text int64 [B, 77])
    -> dict(last_hidden_state [B, 77, width], pooler_output [B, embed_dim]); the whole tower runs in libmesm_b200.so
    (mesm_clip_forward) in fp32 with bf16x3 products - the reference computes the same graph in fp16."""

    def __init__(self, embed_dim, context_length, vocab_size, transformer_width, transformer_heads, transformer_layers):
        super().__init__()
        self.context_length, self.vocab_size, self.embed_dim = context_length, vocab_size, embed_dim
        self.transformer = _ClipTransformer(transformer_width, transformer_layers, transformer_heads)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.initialize_parameters()

    def initialize_parameters(self):            # model/text_encoder.py:297-319
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        w, n = self.transformer.width, self.transformer.layers
        proj_std, attn_std, fc_std = (w ** -0.5) * ((2 * n) ** -0.5), w ** -0.5, (2 * w) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=w ** -0.5)

    @property
    def dtype(self):
        return torch.float32

    def _ctx(self, device):
        lib = _lib.lib()
        sig = tuple((k, v.data_ptr(), v._version) for k, v in self.state_dict().items())
        if getattr(self, "_clip", None) is None or self._clip_sig != sig or self._clip_dev != device:
            if getattr(self, "_clip", None):
                lib.mesm_clip_destroy(self._clip)
            t = self.transformer
            with torch.cuda.device(device):
                c = lib.mesm_clip_create(t.width, t.heads, t.layers, self.context_length, self.vocab_size, self.embed_dim, device.index or 0)
                if not c:
                    raise RuntimeError("mesm_clip_create failed: " + lib.mesm_clip_last_error(None).decode())
                keep = []
                for k, v in self.state_dict().items():
                    w = v.detach().to(device=device, dtype=torch.float32).contiguous()
                    keep.append(w)
                    shape = (ctypes.c_int64 * max(w.dim(), 1))(*w.shape)
                    if lib.mesm_clip_load_weight(c, k.encode(), _ptr(w), shape, w.dim(), 1, _stream()) != 0:
                        raise RuntimeError(lib.mesm_clip_last_error(c).decode())
                if lib.mesm_clip_finalize(c, _stream()) != 0:
                    raise RuntimeError(lib.mesm_clip_last_error(c).decode())
            self.__dict__["_clip"], self.__dict__["_clip_sig"], self.__dict__["_clip_dev"] = c, sig, device
        return self._clip

    def __del__(self):
        try:
            if getattr(self, "_clip", None):
                _lib.lib().mesm_clip_destroy(self._clip)
        except Exception:
            pass

    @torch.no_grad()
    def forward(self, text):
        if not text.is_cuda:
            raise RuntimeError("mesm_b200: `text` must be a CUDA tensor (no CPU fallback)")
        lib = _lib.lib()
        dev = text.device
        c = self._ctx(dev)
        text = text.to(torch.int64).contiguous()
        B = text.shape[0]
        if text.shape[1] != self.context_length:
            raise RuntimeError("CLIPTextEncoder: text must be [B, context_length]")
        x = torch.empty(B, self.context_length, self.transformer.width, dtype=torch.float32, device=dev)
        pooled = torch.empty(B, self.embed_dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = torch.empty(lib.mesm_clip_workspace_bytes(c, B), dtype=torch.uint8, device=dev)
            if lib.mesm_clip_forward(c, _ptr(text), B, _ptr(x), _ptr(pooled), _ptr(ws), ws.numel(), _stream()) != 0:
                raise RuntimeError(lib.mesm_clip_last_error(c).decode())
        return dict(last_hidden_state=x, pooler_output=pooled)


def sample_outclass_neg(num_clips, generator=None):
    """Vectorised equivalent of utils/data_utils.py:113-124: per pair a uniformly random pair of another video group.
    (Same distribution; the reference's per-pair torch.randperm stream is not reproduced — inject ``neg_index`` for
    bit-comparable runs.)"""
    nc = torch.as_tensor(num_clips, dtype=torch.int64, device="cpu")
    B = int(nc.sum())
    end = nc.cumsum(0)
    start = end - nc
    size = torch.repeat_interleave(nc, nc)
    st = torch.repeat_interleave(start, nc)
    if int(size.max()) >= B:
        raise IndexError("sample_outclass_neg needs at least two video groups")
    r = (torch.rand(B, generator=generator) * (B - size).float()).long().clamp_(max=B - 1)
    r = torch.minimum(r, B - size - 1)
    return torch.where(r >= st, r + size, r)


class MESM(_EngineBacked):
    """Drop-in for model.MESM (model/model.py:16-359), inference only.

    forward(video_feat, video_mask, words_id, words_mask, words_weight, num_clips, **kwargs) -> dict with the
    reference's keys (model/model.py:334-351).  Extra kwargs: ``neg_index`` (int64 [B]) replaces the RNG draw of
    model.py:260; ``dataset_name`` selects the SS-MESM grouping branch exactly like the reference."""

    def __init__(self, text_encoder, enhance_encoder, t2v_encoder, transformer, vid_position_embed, txt_position_embed,
                 txt_dim, vid_dim, num_queries, input_dropout, aux_loss=False, max_video_l=75, max_words_l=32,
                 normalize_txt=True, use_txt_pos=False, span_loss_type="l1", n_input_proj=2, rec_fw=False, vocab_size=1111,
                 rec_ss=False, num_recss_layers=2):
        super().__init__()
        if use_txt_pos or not normalize_txt or span_loss_type != "l1" or n_input_proj != 2:
            raise NotImplementedError("only the settings of the shipped configs are implemented "
                                      "(use_txt_pos=False, normalize_txt=True, span_loss_type='l1', n_input_proj=2)")
        if text_encoder is not None and not isinstance(text_encoder, (GloveTextEncoder, CLIPTextEncoder)):
            raise NotImplementedError("text_encoder must be None (word features in `words_id`), a GloveTextEncoder or a CLIPTextEncoder")
        self.text_encoder = text_encoder
        self.enhance_encoder = enhance_encoder
        self.t2v_encoder = t2v_encoder
        self.transformer = transformer
        self.vid_position_embed = vid_position_embed
        self.txt_position_embed = txt_position_embed
        self.num_queries = num_queries
        hidden_dim = transformer.d_model
        self.span_loss_type, self.max_video_l, self.max_words_l = span_loss_type, max_video_l, max_words_l
        self.normalize_txt, self.use_txt_pos, self.n_input_proj = normalize_txt, use_txt_pos, n_input_proj
        self.span_embed = MLP(hidden_dim, hidden_dim, 2, 3)
        self.class_embed = nn.Linear(hidden_dim, 2)
        self.query_embed = nn.Embedding(num_queries, 2)
        relu_args = [True] * 3
        relu_args[n_input_proj - 1] = False
        self.input_txt_proj = nn.Sequential(*[
            LinearLayer(txt_dim, hidden_dim, layer_norm=True, dropout=input_dropout, relu=relu_args[0]),
            LinearLayer(hidden_dim, hidden_dim, layer_norm=True, dropout=input_dropout, relu=relu_args[1])][:n_input_proj])
        self.input_vid_proj = nn.Sequential(*[
            LinearLayer(vid_dim, hidden_dim, layer_norm=True, dropout=input_dropout, relu=relu_args[0]),
            LinearLayer(hidden_dim, hidden_dim, layer_norm=True, dropout=input_dropout, relu=relu_args[1])][:n_input_proj])
        self.saliency_proj1 = nn.Linear(hidden_dim, hidden_dim)
        self.saliency_proj2 = nn.Linear(hidden_dim, hidden_dim)
        self.aux_loss = aux_loss
        self.hidden_dim = hidden_dim
        self.global_rep_token = nn.Parameter(torch.randn(hidden_dim))
        self.global_rep_pos = nn.Parameter(torch.randn(hidden_dim))
        self.rec_fw = rec_fw
        self.txt_dim, self.vid_dim = txt_dim, vid_dim
        if rec_fw:
            num_classes = vocab_size + 1
            self.masked_token = nn.Parameter(torch.zeros(txt_dim).float(), requires_grad=True)
            self.unknown_token = nn.Parameter(torch.zeros(txt_dim).float(), requires_grad=True)
            self.output_txt_proj = nn.Sequential(
                LinearLayer(hidden_dim, hidden_dim, layer_norm=True, dropout=input_dropout, relu=True),
                nn.Linear(hidden_dim, num_classes))
        self.rec_ss = rec_ss
        if rec_ss:
            self.ss_reconstructor = SegSenRecon(
                input_dropout=input_dropout, hidden_dim=hidden_dim, nhead=transformer.nhead, num_layers=num_recss_layers,
                dim_feedforward=transformer.dim_feedforward, dropout=transformer.dropout, activation=transformer.activation,
                normalize_before=transformer.normalize_before)
        self.num_recss_layers = num_recss_layers
        self.chunk_pairs = 0          # 0 = auto (see Engine)
        self._dataset_name = None

    # sub-engines of the child modules are not used by the fused forward: one context holds the whole state_dict
    def _weights_signature(self):
        return (self._dataset_name,) + tuple((p.data_ptr(), p._version) for n, p in self.named_parameters() if not n.startswith("text_encoder."))

    def _engine_cfg(self):
        tr = self.transformer
        return dict(v_feat_dim=self.vid_dim, t_feat_dim=self.txt_dim, hidden_dim=self.hidden_dim, nheads=tr.nhead,
                    dim_feedforward=tr.dim_feedforward, num_queries=self.num_queries,
                    num_recfw_layers=self.enhance_encoder.t2v_encoder.num_layers if self.rec_fw else 0,
                    t2v_layers=self.t2v_encoder.t2v_encoder.num_layers, enc_layers=tr.num_encoder_layers,
                    dec_layers=tr.num_decoder_layers, num_recss_layers=self.num_recss_layers, n_input_proj=self.n_input_proj,
                    rec_fw=self.rec_fw, rec_ss=self.rec_ss, share_MLP=not isinstance(self.enhance_encoder, T2VEncoder_TwoMLP),
                    dataset_name=self._dataset_name, max_words_l=self.max_words_l, max_video_l=self.max_video_l)

    def _engine(self, device):
        eng = getattr(self, "_eng", None)
        if eng is not None and eng.cfg.qvh_grouping != int(self._dataset_name == "qvhighlights"):
            self.__dict__["_eng"] = None          # grouping branch changed: rebuild the context
        return super()._engine(device)

    @torch.no_grad()
    def forward(self, video_feat, video_mask, words_id, words_mask, words_weight, num_clips, **kwargs):
        if kwargs.get("is_training", False) or self.training:
            raise NotImplementedError("mesm_b200.MESM is inference-only (call .eval(); pass is_training=False)")
        name = kwargs.get("dataset_name")
        if self.rec_ss and name not in ("charades", "charades-cg", "charades-cd", "tacos", "qvhighlights"):
            raise KeyError("dataset_name")        # the reference raises KeyError / NotImplementedError here too
        self._dataset_name = "qvhighlights" if name == "qvhighlights" else "charades"
        if isinstance(self.text_encoder, GloveTextEncoder):          # model/model.py:136-143 up to the normalise
            words_feat = self.text_encoder(words_id).masked_fill(words_mask.unsqueeze(-1) == False, 0)  # noqa: E712
        elif isinstance(self.text_encoder, CLIPTextEncoder):         # CLIP_encode_text, model/model.py:103-125 up to the normalise
            words_feat = self.text_encoder(words_id)["last_hidden_state"][:, :self.max_words_l, :]
            words_feat = words_feat.masked_fill(words_mask[:, :self.max_words_l].unsqueeze(-1) == False, 0)  # noqa: E712
        else:
            words_feat = words_id
        eng = self._engine(video_feat.device)
        neg_index = kwargs.get("neg_index")
        if neg_index is None:
            neg_index = sample_outclass_neg(num_clips)
        want = ("core", "rec", "aux") if self.aux_loss else ("core", "rec")
        # `video_len` (host clip counts, added to the batch by mesm_b200.prepare_batch_input) switches the engine to packed
        # variable-length rows; it reaches this forward through **kwargs exactly like the reference's extra batch keys
        o = eng.forward(video_feat, video_mask, words_feat, num_clips, neg_index=neg_index, want=want,
                        video_len=kwargs.get("video_len"), shared_group_video=bool(kwargs.get("shared_group_video", False)))
        out = {"pred_logits": o["pred_logits"], "pred_spans": o["pred_spans"], "saliency_scores": o["saliency_scores"],
               "neg_saliency_scores": o["neg_saliency_scores"]}
        if self.aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_spans": b} for a, b in zip(o["aux_logits"], o["aux_spans"])]
        if self.rec_ss:
            out.update({"projed_video_feat": o["projed_video_feat"], "recon_feat": o["recon_feat"],
                        "projed_recon_feat": o["projed_recon_feat"], "expanded_words_feat": o["expanded_words_feat"],
                        "expanded_words_mask": o["expanded_words_mask"], "enhanced_video_feat": o["enhanced_video_feat"],
                        "projed_words_feat": o["expanded_words_feat"][:, 1:]})
        return out


def build_model(args, vocab=None):
    """runner.build_model (runner.py:255-298) for ``args`` with the reference's attribute names (an argparse
    namespace, or any object / dict with those keys).  text_encoder=None (word features are the input)."""
    g = (lambda k, d=None: args.get(k, d)) if isinstance(args, dict) else (lambda k, d=None: getattr(args, k, d))
    kw = dict(d_model=g("hidden_dim", 256), dropout=g("dropout", 0.1), nhead=g("nheads", 8),
              dim_feedforward=g("dim_feedforward", 1024), normalize_before=g("pre_norm", False), activation="prelu")
    enh_cls = T2VEncoder if g("share_MLP", True) else T2VEncoder_TwoMLP
    enhance = enh_cls(num_encoder_layers=g("num_recfw_layers", 2), **kw)
    t2v = T2VEncoder(num_encoder_layers=g("t2v_layers", 2), **kw)
    transformer = Transformer(num_encoder_layers=g("enc_layers", 2), num_decoder_layers=g("dec_layers", 2),
                              return_intermediate_dec=True, **kw)
    vpos = PositionEmbeddingSine(g("hidden_dim", 256), normalize=True)
    rec_ss = g("rec_ss", True)
    tpos = TrainablePositionalEncoding(g("max_words_l", 32) + 1 if rec_ss else g("max_words_l", 32), g("hidden_dim", 256),
                                       g("input_dropout", 0.5))
    model = MESM(text_encoder=None, enhance_encoder=enhance, t2v_encoder=t2v, transformer=transformer,
                 vid_position_embed=vpos, txt_position_embed=tpos, txt_dim=g("t_feat_dim"), vid_dim=g("v_feat_dim"),
                 num_queries=g("num_queries", 10), input_dropout=g("input_dropout", 0.5), aux_loss=g("aux_loss", True),
                 max_video_l=g("max_video_l", 75), max_words_l=g("max_words_l", 32), normalize_txt=g("normalize_txt", True),
                 use_txt_pos=g("use_txt_pos", False), span_loss_type=g("span_loss_type", "l1"),
                 n_input_proj=g("n_input_proj", 2), rec_fw=g("rec_fw", True), vocab_size=g("vocab_size", 1111),
                 rec_ss=rec_ss, num_recss_layers=g("num_recss_layers", 4))
    return model.eval()
