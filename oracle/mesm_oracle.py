"""CPU restatement (torch fp32, batch-first, functional over a state_dict) of MESM's per-pair inference forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA path and the CPU baseline of bench.py.
Never imported by ``mesm_b200``.

Parity status: PINNED.  ``oracle/gen_golden.py`` (run in the build container, where /root/reference is mounted)
loads the same state_dict into the real ``model.MESM`` and asserts this file reproduces every output and tap to
<= 2e-5 abs for all four shipped shapes incl. ragged batches (mask-coupling quirk) and multi-query video groups;
the reference outputs are committed under tests/golden/ and re-checked by tests/test_oracle_golden.py.

Each function cites the reference lines it restates (paths relative to the reference repo root).
"""
import math

import torch
import torch.nn.functional as F

from .config import OracleConfig

# ----------------------------------------------------------------------------------------------------------------
# matmul precision hook: used only by oracle/precision_study.py to decide the tensor-core operand format
# ("decide with the oracle taps, not by guess").  Default = exact fp32.
# ----------------------------------------------------------------------------------------------------------------
_PREC = {"mode": "fp32", "sites": {}, "site": None}
_MODES = ("fp32", "bf16", "tf32_trunc", "tf32_rn", "bf16x3", "bf16x2", "fp16", "fp16x2a", "fp16x2w", "fp16x3")


def set_matmul_precision(mode: str, sites=None):
    """``mode`` applies to every product; ``sites`` ({"ffn1": mode, "ffn2": mode, ...}) overrides it for the products
    issued under ``_site(name)`` (the per-GEMM precision budget of oracle/precision_study.py)."""
    assert mode in _MODES and all(m in _MODES for m in (sites or {}).values())
    _PREC["mode"] = mode
    _PREC["sites"] = dict(sites or {})


class _site:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.prev = _PREC["site"]
        _PREC["site"] = self.name

    def __exit__(self, *a):
        _PREC["site"] = self.prev


def _tf32(x, rn):
    i = x.contiguous().view(torch.int32)
    if rn:
        i = i + 0x1000
    return (i & ~0x1FFF).view(torch.float32)


def _mm(a, bt):
    """a[..., K] @ bt[..., K]^T-style product under the simulated operand precision (fp32 accumulate)."""
    m = _PREC["sites"].get(_PREC["site"], _PREC["mode"])
    if m == "fp32":
        return a @ bt
    if m.startswith("fp16"):
        ah, bh = a.half().float(), bt.half().float()
        if m == "fp16":
            return ah @ bh
        al, bl = (a - ah).half().float(), (bt - bh).half().float()
        if m == "fp16x2a":
            return ah @ bh + al @ bh          # activations split, weights single
        if m == "fp16x2w":
            return ah @ bh + ah @ bl          # weights split, activations single
        return ah @ bh + (ah @ bl + al @ bh)
    if m == "bf16":
        return a.bfloat16().float() @ bt.bfloat16().float()
    if m in ("tf32_trunc", "tf32_rn"):
        return _tf32(a, m == "tf32_rn") @ _tf32(bt, m == "tf32_rn")
    ah, bh = a.bfloat16().float(), bt.bfloat16().float()
    al, bl = (a - ah).bfloat16().float(), (bt - bh).bfloat16().float()
    if m == "bf16x3":
        return ah @ bh + (ah @ bl + al @ bh)
    return ah @ bh + al @ bh  # bf16x2: activations split, weights single


def linear(x, w, b=None):
    y = _mm(x, w.t())
    return y if b is None else y + b


def layer_norm(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def prelu(x, a):
    return torch.where(x >= 0, x, a * x)


def inverse_sigmoid(x, eps=1e-3):
    """model/transformer.py:36-40 == utils/data_utils.py:139-143."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


# ----------------------------------------------------------------------------------------------------------------
# A1 text post-processing, A2 input projections, A3 positional encoding
# ----------------------------------------------------------------------------------------------------------------
def post_process_text(words_feat):
    """model/model.py:145-152 (text_encoder=None path)."""
    words_feat = F.normalize(words_feat, dim=-1, p=2, eps=1e-5)
    words_mask = words_feat.sum(dim=-1) != 0
    sentence_feat = words_feat.sum(dim=1) / words_mask.sum(dim=1).unsqueeze(-1)
    sentence_feat = F.normalize(sentence_feat, dim=-1, p=2, eps=1e-5)
    return words_feat, words_mask, sentence_feat


def linear_layer(sd, p, x, relu):
    """LinearLayer.forward, model/model.py:427-434 (eval: dropout = identity)."""
    x = layer_norm(x, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"])
    x = linear(x, sd[p + ".net.1.weight"], sd[p + ".net.1.bias"])
    return F.relu(x) if relu else x


def input_proj(sd, name, x, n_input_proj=2):
    """input_{vid,txt}_proj, model/model.py:51-62: relu flags [True]*3 with index n_input_proj-1 False."""
    relu = [True] * 3
    relu[n_input_proj - 1] = False
    for i in range(n_input_proj):
        x = linear_layer(sd, f"{name}.{i}", x, relu[i])
    return x


def position_embedding_sine(mask, num_pos_feats=256, temperature=10000.0):
    """PositionEmbeddingSine.forward with normalize=True, model/position_encoding.py:51-72."""
    x_embed = mask.cumsum(1, dtype=torch.float32)
    x_embed = x_embed / (x_embed[:, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / num_pos_feats)
    pos = x_embed[:, :, None] / dim_t
    return torch.stack((pos[:, :, 0::2].sin(), pos[:, :, 1::2].cos()), dim=3).flatten(2)


# ----------------------------------------------------------------------------------------------------------------
# attention cores
# ----------------------------------------------------------------------------------------------------------------
def _heads(x, H):
    B, L, E = x.shape
    return x.view(B, L, H, E // H).permute(0, 2, 1, 3)          # [B,H,L,hd]


def _attend(q, k, v, masked, H):
    """q [B,Lq,E] (already scaled), k [B,Lk,E], v [B,Lk,Ev]; masked bool broadcastable to [B,H,Lq,Lk] (True = -inf).
    Softmax of (scores - rowmax) as model/attention.py:360-384 (== torch's softmax)."""
    with _site(_PREC["site"] or "attn"):
        s = _mm(_heads(q, H), _heads(k, H).transpose(-1, -2))
        s = s.masked_fill(masked, float("-inf"))
        p = torch.softmax(s, dim=-1)
        o = _mm(p, _heads(v, H))
    B, _, Lq, hv = o.shape
    return o.permute(0, 2, 1, 3).reshape(B, Lq, H * hv)


def packed_mha(sd, p, q_in, k_in, v_in, masked, H):
    """nn.MultiheadAttention(d, H) as called at model/transformer.py:532-533, 643-644: packed in_proj rows
    [Wq;Wk;Wv], q scaled by head_dim^-0.5, bool masks -> -inf, out_proj."""
    d = q_in.shape[-1]
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = linear(q_in, w[:d], b[:d]) * (float(d // H) ** -0.5)
    k = linear(k_in, w[d:2 * d], b[d:2 * d])
    v = linear(v_in, w[2 * d:], b[2 * d:])
    o = _attend(q, k, v, masked, H)
    return linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def t2v_mask(q_pad, k_pad, H):
    """The mask the reference *actually* applies in T2V_TransformerEncoderLayer.forward_post
    (model/transformer.py:528-533): attn_mask = (q_pad (x) k_pad).bool().repeat(H,1,1) is indexed by torch as
    [b*H + h], i.e. pair b / head h sees the pattern of pair b' = (b*H+h) % B; it is OR-ed with key_padding_mask.
    masked[b,h,q,k] = k_pad[b,k] | (q_pad[b',q] & k_pad[b',k])."""
    B = q_pad.shape[0]
    bp = (torch.arange(B)[:, None] * H + torch.arange(H)[None, :]) % B          # [B,H]
    cross = q_pad[bp][:, :, :, None] & k_pad[bp][:, :, None, :]                   # [B,H,Lq,Lk]
    return cross | k_pad[:, None, None, :]


def ffn_post(sd, p, x, suffix=""):
    """x -> LN2(x + W2 PReLU(W1 LN1(x)))   (T2V layer tail, model/transformer.py:536-539)."""
    y = layer_norm(x, sd[p + f"norm1{suffix}.weight"], sd[p + f"norm1{suffix}.bias"])
    with _site("ffn1"):
        h = prelu(linear(y, sd[p + f"linear1{suffix}.weight"], sd[p + f"linear1{suffix}.bias"]), sd[p + "activation.weight"])
    with _site("ffn2"):
        y = linear(h, sd[p + f"linear2{suffix}.weight"], sd[p + f"linear2{suffix}.bias"])
    return layer_norm(x + y, sd[p + f"norm2{suffix}.weight"], sd[p + f"norm2{suffix}.bias"])


def t2v_layer(sd, p, txt, vid, txt_pad, pos_txt, vid_pad, pos_vid, H):
    """T2V_TransformerEncoderLayer.forward_post, model/transformer.py:508-540 (and the TwoMLP variant with
    is_MLM=False, 573-612): cross-attention q = vid(+pos), k = txt(+pos), v = txt; post-norm FFN with PReLU."""
    q_in = vid if pos_vid is None else vid + pos_vid
    k_in = txt if pos_txt is None else txt + pos_txt
    a = packed_mha(sd, p + "self_attn.", q_in, k_in, txt, t2v_mask(vid_pad, txt_pad, H), H)
    return ffn_post(sd, p, vid + a)


def t2v_stack(sd, p, n_layers, txt, vid, txt_pad, pos_txt, vid_pad, pos_vid, H):
    """T2V_TransformerEncoder.forward, model/transformer.py:216-242 (norm=None: pre_norm is false everywhere)."""
    for i in range(n_layers):
        vid = t2v_layer(sd, f"{p}.layers.{i}.", txt, vid, txt_pad, pos_txt, vid_pad, pos_vid, H)
    return vid


def encoder_layer(sd, p, src, pad, pos, H):
    """TransformerEncoderLayer.forward_post, model/transformer.py:637-650."""
    qk = src + pos
    a = packed_mha(sd, p + "self_attn.", qk, qk, src, pad[:, None, None, :], H)
    src = layer_norm(src + a, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    with _site("ffn1"):
        h = prelu(linear(src, sd[p + "linear1.weight"], sd[p + "linear1.bias"]), sd[p + "activation.weight"])
    with _site("ffn2"):
        y = linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return layer_norm(src + y, sd[p + "norm2.weight"], sd[p + "norm2.bias"])


def mlp(sd, p, x, n):
    """MLP.forward, model/transformer.py:30-33 / model/model.py:406-409."""
    for i in range(n):
        x = linear(x, sd[f"{p}.layers.{i}.weight"], sd[f"{p}.layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x


def gen_sineembed_for_position(ref, dim=256):
    """model/transformer.py:43-59 on batch-first ref [B,nq,2]."""
    each = dim // 2
    dim_t = torch.arange(each, dtype=torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / each)
    outs = []
    for c in range(2):
        e = (ref[..., c] * (2 * math.pi))[..., None] / dim_t
        outs.append(torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-2))
    return torch.cat(outs, dim=-1)


def plain_mha(sd, p, q, k, v, pad, H):
    """model/attention.py:185-394 (no in-projection): scale by (E/H)^-0.5, key padding -> -inf, softmax(x - max),
    PV, out_proj(vdim, vdim)."""
    E = q.shape[-1]
    o = _attend(q * (float(E // H) ** -0.5), k, v,
                pad[:, None, None, :] if pad is not None else torch.zeros(1, 1, 1, 1, dtype=torch.bool), H)
    return linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def decoder_layer(sd, p, tgt, memory, mem_pad, pos, query_pos, query_sine, is_first, H):
    """TransformerDecoderLayer.forward, model/transformer.py:723-797 (batch-first)."""
    def L(n, x):
        return linear(x, sd[p + n + ".weight"], sd[p + n + ".bias"])
    q = L("sa_qcontent_proj", tgt) + L("sa_qpos_proj", query_pos)
    k = L("sa_kcontent_proj", tgt) + L("sa_kpos_proj", query_pos)
    v = L("sa_v_proj", tgt)
    tgt = layer_norm(tgt + plain_mha(sd, p + "self_attn.", q, k, v, None, H), sd[p + "norm1.weight"], sd[p + "norm1.bias"])

    q = L("ca_qcontent_proj", tgt)
    k = L("ca_kcontent_proj", memory)
    v = L("ca_v_proj", memory)
    k_pos = L("ca_kpos_proj", pos)
    if is_first:
        q = q + L("ca_qpos_proj", query_pos)
        k = k + k_pos
    B, nq, d = q.shape
    hw = k.shape[1]
    hd = d // H
    sine = L("ca_qpos_sine_proj", query_sine)
    q = torch.cat([q.view(B, nq, H, hd), sine.view(B, nq, H, hd)], dim=3).view(B, nq, 2 * d)
    k = torch.cat([k.view(B, hw, H, hd), k_pos.view(B, hw, H, hd)], dim=3).view(B, hw, 2 * d)
    tgt = layer_norm(tgt + plain_mha(sd, p + "cross_attn.", q, k, v, mem_pad, H), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    y = linear(prelu(L("linear1", tgt), sd[p + "activation.weight"]), sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return layer_norm(tgt + y, sd[p + "norm3.weight"], sd[p + "norm3.bias"])


def decoder(sd, cfg, memory, mem_pad, pos, query_embed):
    """TransformerDecoder.forward, model/transformer.py:333-420 (batch-first; tgt = 0, :201)."""
    p = "transformer.decoder."
    B = memory.shape[0]
    nq = query_embed.shape[0]
    out = torch.zeros(B, nq, cfg.hidden_dim)
    ref = query_embed.sigmoid()[None].expand(B, nq, 2)
    refs, inter = [ref], []
    for lid in range(cfg.dec_layers):
        sine = gen_sineembed_for_position(ref, cfg.hidden_dim)
        query_pos = mlp(sd, p + "ref_point_head", sine, 2)
        if lid > 0:
            sine = sine * mlp(sd, p + "query_scale", out, 2)
        reft = mlp(sd, p + "ref_anchor_head", out, 2).sigmoid()
        sine = sine * (reft[..., 0] / ref[..., 1]).unsqueeze(-1)
        out = decoder_layer(sd, f"{p}layers.{lid}.", out, memory, mem_pad, pos, query_pos, sine, lid == 0, cfg.nheads)
        tmp = mlp(sd, p + "bbox_embed", out, 3)
        new_ref = (tmp[..., :2] + inverse_sigmoid(ref)).sigmoid()
        if lid != cfg.dec_layers - 1:
            refs.append(new_ref)
        ref = new_ref
        inter.append(layer_norm(out, sd[p + "norm.weight"], sd[p + "norm.bias"]))
    return torch.stack(inter), torch.stack(refs)          # [nl,B,nq,d], [nl,B,nq,2]


def transformer(sd, cfg, src, pad, query_embed, pos, with_decoder=True):
    """Transformer.forward, model/transformer.py:174-205: global token prepended with key-padding flag True
    (185-186: it queries but is never a key), encoder, then DAB-DETR decoder on the local memory."""
    B = src.shape[0]
    g = sd["global_rep_token"].view(1, 1, -1).expand(B, 1, -1)
    gp = sd["global_rep_pos"].view(1, 1, -1).expand(B, 1, -1)
    pad_g = torch.cat([torch.ones(B, 1, dtype=torch.bool), pad], dim=1)
    x = torch.cat([g, src], dim=1)
    pos_g = torch.cat([gp, pos], dim=1)
    for i in range(cfg.enc_layers):
        x = encoder_layer(sd, f"transformer.encoder.layers.{i}.", x, pad_g, pos_g, cfg.nheads)
    memory_global, memory = x[:, 0], x[:, 1:]
    if not with_decoder:
        return None, None, memory, memory_global
    hs, refs = decoder(sd, cfg, memory, pad, pos, query_embed)
    return hs, refs, memory, memory_global


# ----------------------------------------------------------------------------------------------------------------
# SS-MESM sentence reconstructor (A5)
# ----------------------------------------------------------------------------------------------------------------
def _group_of(num_clips):
    nc = num_clips.tolist()
    grp = torch.repeat_interleave(torch.arange(len(nc)), num_clips)
    idx_in_grp = torch.cat([torch.arange(n) for n in nc])
    return grp, idx_in_grp


def ss_reconstruct(sd, cfg, video_feat, video_mask, sentence_feat, num_clips):
    """model/model.py:184-207 + SegSenRecon.forward 467-503: per pair, one masked sentence slot queries the
    projected clips of its video group through 4 T2V layers (no positional terms, :482)."""
    B = video_feat.shape[0]
    nc = num_clips.tolist()
    grp, slot = _group_of(num_clips)
    if cfg.dataset_name in ("charades", "charades-cg", "charades-cd", "tacos"):
        bvid, bmask = video_feat, video_mask                                     # :186-189
    elif cfg.dataset_name == "qvhighlights":                                     # :191-195
        per_group = torch.split(video_mask, nc)
        glen = [int(m.sum()) for m in per_group]
        rows = torch.split(video_feat[video_mask], glen)
        Lg = max(glen)
        bvid = torch.zeros(B, Lg, video_feat.shape[-1])
        bmask = torch.zeros(B, Lg, dtype=torch.bool)
        for b in range(B):
            g = int(grp[b])
            bvid[b, :glen[g]] = rows[g]
            bmask[b, :glen[g]] = True
    else:
        raise NotImplementedError(cfg.dataset_name)
    max_nc = max(nc)
    bsent = torch.zeros(B, max_nc, sentence_feat.shape[-1])                      # split_expand_and_pad, :199
    smask = torch.zeros(B, max_nc, dtype=torch.bool)
    start = 0
    for g, n in enumerate(nc):
        for i in range(n):
            bsent[start + i, :n] = sentence_feat[start:start + n]
            smask[start + i, :n] = True
        start += n
    bvid = input_proj(sd, "input_vid_proj", bvid, cfg.n_input_proj)              # :201
    bsent = input_proj(sd, "input_txt_proj", bsent, cfg.n_input_proj)            # :202
    bsent[torch.arange(B), slot] = sd["ss_reconstructor.masked_sent_token"]      # _sequence_mask_sent, :490-503
    out = t2v_stack(sd, "ss_reconstructor.recon_trans", cfg.num_recss_layers, bvid, bsent, ~bmask, None, ~smask, None,
                    cfg.nheads)
    recon_feat = F.normalize(out[torch.arange(B), slot])                         # :486 (dim=1, eps 1e-12)
    p = "ss_reconstructor.output_sent_proj"
    proj = linear_layer(sd, p + ".1", linear_layer(sd, p + ".0", recon_feat, True), False)
    return recon_feat, proj


# ----------------------------------------------------------------------------------------------------------------
# full forward (model/model.py:154-359, eval mode) and alignment scores (model/criterion.py:241-266)
# ----------------------------------------------------------------------------------------------------------------
def _saliency(sd, memory, memory_global, d):
    a = linear(memory, sd["saliency_proj1.weight"], sd["saliency_proj1.bias"])
    b = linear(memory_global, sd["saliency_proj2.weight"], sd["saliency_proj2.bias"])
    return (a * b.unsqueeze(1)).sum(-1) / math.sqrt(d)                            # model/model.py:301


@torch.no_grad()
def mesm_forward(sd, cfg: OracleConfig, video_feat, video_mask, words_feat, num_clips, neg_index=None,
                 with_neg=True):
    """Returns the reference's output dict (model/model.py:334-351) plus intermediate taps.

    ``words_feat`` is the [B,Lt,Dt] word-feature tensor the reference receives as ``words_id`` when
    ``text_encoder is None`` (model/model.py:160-161).  ``neg_index`` replaces the RNG draw of
    ``sample_outclass_neg`` (model/model.py:260)."""
    H, d = cfg.nheads, cfg.hidden_dim
    video_feat, words_feat = video_feat.float(), words_feat.float()
    words_feat, words_mask, sentence_feat = post_process_text(words_feat)
    projed_video = input_proj(sd, "input_vid_proj", video_feat, cfg.n_input_proj)
    projed_words = input_proj(sd, "input_txt_proj", words_feat, cfg.n_input_proj)
    vid_pos = position_embedding_sine(video_mask, d)
    vpad = ~video_mask

    def enhance(words, wmask):
        if not cfg.rec_fw:
            return projed_video
        return t2v_stack(sd, "enhance_encoder.t2v_encoder", cfg.num_recfw_layers, words, projed_video, ~wmask,
                         None, vpad, vid_pos, H)                                  # txt_position = 0 (:169-172)

    enhanced = enhance(projed_words, words_mask)
    out = {}
    if cfg.rec_ss:
        recon_feat, projed_recon = ss_reconstruct(sd, cfg, video_feat, video_mask, sentence_feat, num_clips)
        exp_words = torch.cat([recon_feat.unsqueeze(1), projed_words], dim=1)     # :217-219
        exp_mask = torch.cat([torch.ones(len(words_mask), 1, dtype=torch.bool), words_mask], dim=1)
    else:
        exp_words, exp_mask = projed_words, words_mask

    def align_and_detr(ewords, emask, enh, with_decoder):
        enc = t2v_stack(sd, "t2v_encoder.t2v_encoder", cfg.t2v_layers, ewords, enh, ~emask, None, vpad, vid_pos, H)
        return (enc,) + tuple(transformer(sd, cfg, enc, vpad, sd["query_embed.weight"], vid_pos, with_decoder))

    encoded, hs, refs, memory, memory_global = align_and_detr(exp_words, exp_mask, enhanced, True)
    outputs_class = linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])  # :246
    outputs_coord = (mlp(sd, "span_embed", hs, 3) + inverse_sigmoid(refs)).sigmoid()  # :247-252
    out.update(pred_logits=outputs_class[-1], pred_spans=outputs_coord[-1],
               saliency_scores=_saliency(sd, memory, memory_global, d))
    if with_neg:
        assert neg_index is not None
        n_exp, n_mask = exp_words[neg_index], exp_mask[neg_index]                 # :261-271
        n_words, n_wmask = (n_exp[:, 1:], n_mask[:, 1:]) if cfg.rec_ss else (n_exp, n_mask)
        _, _, _, n_mem, n_glob = align_and_detr(n_exp, n_mask, enhance(n_words, n_wmask), False)
        out["neg_saliency_scores"] = _saliency(sd, n_mem, n_glob, d)
    if cfg.aux_loss:
        out["aux_outputs"] = [{"pred_logits": a, "pred_spans": b} for a, b in zip(outputs_class[:-1], outputs_coord[:-1])]
    if cfg.rec_ss:
        out.update(projed_video_feat=projed_video, recon_feat=recon_feat, projed_recon_feat=projed_recon,
                   expanded_words_feat=exp_words, expanded_words_mask=exp_mask, enhanced_video_feat=enhanced,
                   projed_words_feat=projed_words)
    out["_taps"] = dict(words_feat=words_feat, words_mask=words_mask, sentence_feat=sentence_feat, vid_pos=vid_pos,
                        encoded_video_feat=encoded, memory=memory, memory_global=memory_global, hs=hs,
                        references=refs)
    return out


@torch.no_grad()
def align_scores(projed_video_feat, clip_mask, expanded_words_feat, expanded_words_mask, tau=0.5):
    """Segment-sentence alignment scores, model/criterion.py:241-266: masked means, L2-normalise (eps 1e-12),
    cos_sim / tau.  Returns S[B,B] (rows: clips of pair i, cols: words of pair j)."""
    cm = clip_mask.unsqueeze(-1)
    clip = (projed_video_feat * cm).sum(1) / cm.sum(1)
    wm = expanded_words_mask.unsqueeze(-1)
    words = (expanded_words_feat * wm).sum(1) / wm.sum(1)
    return F.normalize(clip, dim=-1, p=2) @ F.normalize(words, dim=-1, p=2).t() / tau
