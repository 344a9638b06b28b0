"""Deterministic random weights under the reference's ``state_dict`` key names, and seeded synthetic inputs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``state_spec`` restates, key by key, what ``model.MESM(...).state_dict()`` holds when wired as
``runner.py:255-298`` does with ``text_encoder=None`` (module tree: model/model.py:29-98, 412-465;
model/transformer.py:62-72, 108-116, 139-152, 282-331, 485-503, 562-571, 615-632, 676-718;
model/attention.py:87-100; model/position_encoding.py:13-17).  ``oracle/gen_golden.py`` checks it against the
real reference module (key set and shapes) in the build container.

Values are NOT the reference's init: every tensor is drawn from a per-key seeded CPU generator so that the same
numbers are reproducible on any box without shipping 55 MB of weights, and so that tensors the reference
initialises to a constant (LayerNorm affine, PReLU slope, out_proj bias, bbox_embed last layer, masked tokens)
still exercise their code path.
"""
import hashlib
import math

import torch

from .config import OracleConfig


def _attn_ffn_layer(p, d, ff, packed_in_proj=True):
    spec = []
    if packed_in_proj:
        spec += [(p + "self_attn.in_proj_weight", (3 * d, d)), (p + "self_attn.in_proj_bias", (3 * d,))]
    spec += [(p + "self_attn.out_proj.weight", (d, d)), (p + "self_attn.out_proj.bias", (d,)),
             (p + "linear1.weight", (ff, d)), (p + "linear1.bias", (ff,)),
             (p + "linear2.weight", (d, ff)), (p + "linear2.bias", (d,)),
             (p + "norm1.weight", (d,)), (p + "norm1.bias", (d,)),
             (p + "norm2.weight", (d,)), (p + "norm2.bias", (d,)),
             (p + "activation.weight", (1,))]
    return spec


def _linear(p, out_f, in_f):
    return [(p + ".weight", (out_f, in_f)), (p + ".bias", (out_f,))]


def _mlp(p, dims):
    spec = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        spec += _linear(f"{p}.layers.{i}", b, a)
    return spec


def _linear_layer(p, in_f, out_f):
    return [(p + ".LayerNorm.weight", (in_f,)), (p + ".LayerNorm.bias", (in_f,))] + _linear(p + ".net.1", out_f, in_f)


def state_spec(cfg: OracleConfig):
    """[(key, shape)] of the reference MESM state_dict (text_encoder=None)."""
    d, ff = cfg.hidden_dim, cfg.dim_feedforward
    spec = []
    for i in range(cfg.num_recfw_layers):
        p = f"enhance_encoder.t2v_encoder.layers.{i}."
        spec += _attn_ffn_layer(p, d, ff)
        if not cfg.share_mlp:  # T2V_TransformerEncoderLayer_TwoMLP, transformer.py:562-571
            spec += _linear(p + "linear1_1", ff, d) + _linear(p + "linear2_1", d, ff)
            spec += [(p + "norm1_1.weight", (d,)), (p + "norm1_1.bias", (d,)),
                     (p + "norm2_1.weight", (d,)), (p + "norm2_1.bias", (d,))]
    for i in range(cfg.t2v_layers):
        spec += _attn_ffn_layer(f"t2v_encoder.t2v_encoder.layers.{i}.", d, ff)
    for i in range(cfg.enc_layers):
        spec += _attn_ffn_layer(f"transformer.encoder.layers.{i}.", d, ff)
    for i in range(cfg.dec_layers):
        p = f"transformer.decoder.layers.{i}."
        for n in ("sa_qcontent_proj", "sa_qpos_proj", "sa_kcontent_proj", "sa_kpos_proj", "sa_v_proj",
                  "ca_qcontent_proj", "ca_kcontent_proj", "ca_kpos_proj", "ca_v_proj", "ca_qpos_sine_proj"):
            spec += _linear(p + n, d, d)
        if i == 0:  # ca_qpos_proj = None for layers >= 1, transformer.py:329-331
            spec += _linear(p + "ca_qpos_proj", d, d)
        spec += _linear(p + "self_attn.out_proj", d, d) + _linear(p + "cross_attn.out_proj", d, d)
        spec += _linear(p + "linear1", ff, d) + _linear(p + "linear2", d, ff)
        for n in ("norm1", "norm2", "norm3"):
            spec += [(p + n + ".weight", (d,)), (p + n + ".bias", (d,))]
        spec += [(p + "activation.weight", (1,))]
    spec += [("transformer.decoder.norm.weight", (d,)), ("transformer.decoder.norm.bias", (d,))]
    spec += _mlp("transformer.decoder.query_scale", [d, d, d])
    spec += _mlp("transformer.decoder.ref_point_head", [d, d, d])
    spec += _mlp("transformer.decoder.bbox_embed", [d, d, d, 2])
    spec += _mlp("transformer.decoder.ref_anchor_head", [d, d, 1])
    n_pos = cfg.max_words_l + 1 if cfg.rec_ss else cfg.max_words_l
    spec += [("txt_position_embed.position_embeddings.weight", (n_pos, d)),
             ("txt_position_embed.LayerNorm.weight", (d,)), ("txt_position_embed.LayerNorm.bias", (d,))]
    spec += _mlp("span_embed", [d, d, d, 2])
    spec += _linear("class_embed", 2, d)
    spec += [("query_embed.weight", (cfg.num_queries, 2))]
    dims_t = [cfg.t_feat_dim] + [d] * cfg.n_input_proj
    dims_v = [cfg.v_feat_dim] + [d] * cfg.n_input_proj
    for i in range(cfg.n_input_proj):
        spec += _linear_layer(f"input_txt_proj.{i}", dims_t[i], dims_t[i + 1])
        spec += _linear_layer(f"input_vid_proj.{i}", dims_v[i], dims_v[i + 1])
    spec += _linear("saliency_proj1", d, d) + _linear("saliency_proj2", d, d)
    spec += [("global_rep_token", (d,)), ("global_rep_pos", (d,))]
    if cfg.rec_fw:
        spec += [("masked_token", (cfg.t_feat_dim,)), ("unknown_token", (cfg.t_feat_dim,))]
        spec += _linear_layer("output_txt_proj.0", d, d) + _linear("output_txt_proj.1", cfg.vocab_size + 1, d)
    if cfg.rec_ss:
        spec += [("ss_reconstructor.masked_sent_token", (d,))]
        for i in range(cfg.num_recss_layers):
            spec += _attn_ffn_layer(f"ss_reconstructor.recon_trans.layers.{i}.", d, ff)
        spec += _linear_layer("ss_reconstructor.output_sent_proj.0", d, d)
        spec += _linear_layer("ss_reconstructor.output_sent_proj.1", d, d)
    return spec


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def make_state_dict(cfg: OracleConfig, seed: int = 0):
    """Deterministic fp32 CPU state_dict with the reference key names (see module docstring)."""
    sd = {}
    for key, shape in state_spec(cfg):
        g = torch.Generator().manual_seed(_key_seed(seed, key))
        leaf = key.rsplit(".", 1)[-1]
        is_norm = ("norm" in key.lower()) and leaf in ("weight", "bias") and len(shape) == 1
        if key.endswith("activation.weight"):
            t = 0.25 + 0.05 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 2 and key != "query_embed.weight":
            fan_out, fan_in = shape
            a = math.sqrt(6.0 / (fan_in + fan_out))          # xavier-uniform bound (transformer.py:78-81)
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif key == "query_embed.weight":
            t = torch.randn(shape, generator=g)
        elif key in ("global_rep_token", "global_rep_pos"):
            t = torch.randn(shape, generator=g)
        elif key.endswith("masked_sent_token") or key in ("masked_token", "unknown_token"):
            t = 0.5 * torch.randn(shape, generator=g)
        else:  # biases
            t = 0.05 * torch.randn(shape, generator=g)
        sd[key] = t.to(torch.float32).contiguous()
    return sd


def make_inputs(cfg: OracleConfig, num_clips, seed: int = 0, lv=None, lt=None, ragged_video=True, ragged_text=True,
                min_video_frac=0.5, min_words=3, dur_range=(10.0, 150.0), f16_features=False):
    """Seeded synthetic batch shaped like ``collate``'s output (dataset/base.py:288-355).

    num_clips: list of queries per video group; B = sum.  All queries of a group share the video
    (dataset/base.py:307-312).  Video features: per-source L2-normalised N(0,1) blocks + tef columns
    (dataset/base.py:225-230); word features N(0,1), zero rows past each query's length.
    Returns dict(video_feat f32[B,Lv,Dv], video_mask bool[B,Lv], words_feat f32[B,Lt,Dt], num_clips i64[G],
                 duration f32[B], words_len i64[B], video_len i64[B]).
    """
    g = torch.Generator().manual_seed(_key_seed(seed, "inputs:" + cfg.name))
    lv = cfg.max_video_l if lv is None else lv
    lt = cfg.max_words_l if lt is None else lt
    G, B = len(num_clips), int(sum(num_clips))
    dv = cfg.v_feat_dim - 2
    split = 512 if dv > 512 and cfg.t_feat_dim == 512 else dv          # CLIP(512)+SlowFast(2304) | single source
    vids, vlens = [], []
    for gi in range(G):
        L = lv if not ragged_video else int(torch.randint(max(1, int(lv * min_video_frac)), lv + 1, (1,), generator=g))
        if gi == 0:
            L = lv                                                          # keep the padded length = lv
        x = torch.randn(L, dv, generator=g)
        x[:, :split] = torch.nn.functional.normalize(x[:, :split], dim=1, eps=1e-5)
        if split < dv:
            x[:, split:] = torch.nn.functional.normalize(x[:, split:], dim=1, eps=1e-5)
        st = torch.arange(0, L, 1.0) / L
        x = torch.cat([x, torch.stack([st, st + 1.0 / L], dim=1)], dim=1)
        vids.append(x)
        vlens.append(L)
    video_feat = torch.zeros(B, lv, cfg.v_feat_dim)
    video_mask = torch.zeros(B, lv, dtype=torch.bool)
    video_len = torch.zeros(B, dtype=torch.int64)
    b = 0
    for gi, n in enumerate(num_clips):
        for _ in range(n):
            video_feat[b, :vlens[gi]] = vids[gi]
            video_mask[b, :vlens[gi]] = True
            video_len[b] = vlens[gi]
            b += 1
    if f16_features:                 # the 16-bit feature-storage option: the model is fed the fp16-representable values
        video_feat = video_feat.half().float()
    words = torch.randn(B, lt, cfg.t_feat_dim, generator=g)
    wl = torch.randint(min(min_words, lt), lt + 1, (B,), generator=g) if ragged_text else torch.full((B,), lt)
    wl[0] = lt
    for b in range(B):
        words[b, wl[b]:] = 0
    duration = torch.rand(G, generator=g) * (dur_range[1] - dur_range[0]) + dur_range[0]
    duration = torch.repeat_interleave(duration, torch.tensor(num_clips))
    return dict(video_feat=video_feat, video_mask=video_mask, words_feat=words,
                num_clips=torch.tensor(num_clips, dtype=torch.int64), duration=duration.to(torch.float32),
                words_len=wl.to(torch.int64), video_len=video_len)


def make_neg_index(num_clips, seed: int = 0):
    """Deterministic stand-in for ``sample_outclass_neg`` (utils/data_utils.py:113-124): for every pair, one pair
    index drawn from a different video group.  Needs >= 2 groups (the reference raises otherwise)."""
    g = torch.Generator().manual_seed(_key_seed(seed, "neg_index"))
    nc = torch.as_tensor(num_clips, dtype=torch.int64)
    end = nc.cumsum(0)
    start = end - nc
    B = int(end[-1])
    out = []
    for gi in range(len(nc)):
        for _ in range(int(nc[gi])):
            cand = torch.cat([torch.arange(0, int(start[gi])), torch.arange(int(end[gi]), B)])
            out.append(cand[int(torch.randint(0, len(cand), (1,), generator=g))])
    return torch.stack(out)


# raw per-source clip features of the feature-ingest front-end fixtures (tests/golden/frontend.npz): name -> (clip counts per
# source, feature dims per source, max_video_l, stored as fp16)
FRONTEND_CASES = {
    "csf_short": ((21, 23), (512, 2304), 75, False),        # shipped C+SF dims, shorter than max_video_l: no pooling
    "csf_pool": ((160, 157), (64, 200), 75, True),          # mean-pool down-sampling (dataset/base.py:100-114), 16-bit storage
    "vgg_pool": ((333,), (300,), 200, True),
    "c3d_exact": ((200,), (128,), 200, False),              # length == max_video_l
}


def make_raw_features(name):
    """Seeded raw (un-normalised) per-source features of a front-end fixture: list of [L_s, D_s] tensors (fp16 or fp32)."""
    lens, dims, max_l, f16 = FRONTEND_CASES[name]
    g = torch.Generator().manual_seed(_key_seed(7, "frontend:" + name))
    raws = [torch.randn(n, d, generator=g) * (1 + i) for i, (n, d) in enumerate(zip(lens, dims))]
    return [r.half() if f16 else r for r in raws], max_l


CLIP_TEXT_CFG = dict(embed_dim=512, context_length=77, vocab_size=1000, transformer_width=512, transformer_heads=8, transformer_layers=12)


def make_clip_state_dict(seed=0, cfg=None):
    """Seeded state_dict of a CLIPTextEncoder (model/text_encoder.py:240-354; the shipped ViT-B/32 text tower shape with a small
    vocabulary), per-key generators like make_state_dict; std per tensor kind as initialize_parameters (:297-319) so that the
    activations have realistic magnitudes."""
    c = cfg or CLIP_TEXT_CFG
    W, n = c["transformer_width"], c["transformer_layers"]
    spec = [("token_embedding.weight", (c["vocab_size"], W), 0.02), ("positional_embedding", (c["context_length"], W), 0.01),
            ("ln_final.weight", (W,), None), ("ln_final.bias", (W,), 0.02), ("text_projection", (W, c["embed_dim"]), W ** -0.5)]
    for l in range(n):
        p = f"transformer.resblocks.{l}."
        spec += [(p + "attn.in_proj_weight", (3 * W, W), W ** -0.5), (p + "attn.in_proj_bias", (3 * W,), 0.02),
                 (p + "attn.out_proj.weight", (W, W), (W ** -0.5) * ((2 * n) ** -0.5)), (p + "attn.out_proj.bias", (W,), 0.02),
                 (p + "ln_1.weight", (W,), None), (p + "ln_1.bias", (W,), 0.02), (p + "ln_2.weight", (W,), None), (p + "ln_2.bias", (W,), 0.02),
                 (p + "mlp.c_fc.weight", (4 * W, W), (2 * W) ** -0.5), (p + "mlp.c_fc.bias", (4 * W,), 0.02),
                 (p + "mlp.c_proj.weight", (W, 4 * W), (W ** -0.5) * ((2 * n) ** -0.5)), (p + "mlp.c_proj.bias", (W,), 0.02)]
    sd = {}
    for k, shape, std in spec:
        g = torch.Generator().manual_seed(_key_seed(seed, "clip:" + k))
        sd[k] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if std is None else torch.randn(shape, generator=g) * std
    return sd


def make_clip_tokens(B=5, seed=0, cfg=None):
    """Token ids [B, 77]: sot, words, eot = the largest id of the row (text.argmax picks it, :346), zero padding."""
    c = cfg or CLIP_TEXT_CFG
    g = torch.Generator().manual_seed(_key_seed(seed, "clip:tokens"))
    V, L = c["vocab_size"], c["context_length"]
    text = torch.zeros(B, L, dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(3, 31, (1,), generator=g))
        text[b, 0] = V - 2
        text[b, 1:n + 1] = torch.randint(1, V - 2, (n,), generator=g)
        text[b, n + 1] = V - 1
    return text
