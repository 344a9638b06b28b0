#!/usr/bin/env python
"""bench.py — video-query pairs/sec of the MESM per-pair inference path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--config NAME]     our CUDA path (one process per GPU; torchrun for N>1)
    python bench.py --impl reference [...]                              the reference's own modules on the host CPU cores

A step = one pass of the whole hot path - MESM.forward incl. the negative branch, the segment-sentence alignment scores
(model/criterion.py:241-266) and span decode / post-processing / NMS - over one batch of synthetic pairs.  Default workload =
BASELINE.json configs[1]: Charades-STA C+SF shape, 4096 pairs per GPU, inputs resident in HBM (far larger than L2).
``--config`` selects the other BASELINE configs (charades_vgg = configs[2], tacos = configs[3] incl. the 100-candidate NMS pass,
qvh = the shape of configs[0] / [4]; the 1 M-pair sweep itself is sweep.py).  `e2e` is the same work through the public Python
API with the inputs starting in pinned HOST memory (H2D inside the timed region) and the ranked windows read back to the host.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import dataclasses
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_COMMON = dict(hidden_dim=256, nheads=8, dim_feedforward=1024, num_queries=10, num_recfw_layers=2, t2v_layers=2, enc_layers=2,
               dec_layers=2, num_recss_layers=4, n_input_proj=2, rec_fw=True, rec_ss=True, aux_loss=True, vocab_size=1111)
# model keys = the reference's JSON configs (config/*/*.json); "wl" = the synthetic workload of SURVEY 8d (C1..C4)
BENCH_CONFIGS = {
    "charades_csf": dict(_COMMON, dataset_name="charades", v_feat_dim=2818, t_feat_dim=512, share_MLP=True, max_words_l=16,
                         max_video_l=194, clip_len=1.0, max_ts_val=150.0,
                         wl=dict(name="charades_sta_c+sf_inference_batch4096", baseline="configs[1]", oracle="charades_csf", split=512,
                                 group_max=4, ragged=True, dur=(10.0, 60.0), dense_nms=0,
                                 algo_flops=2.51e9, algo_bytes=2219746)),
    "charades_vgg": dict(_COMMON, dataset_name="charades", v_feat_dim=4098, t_feat_dim=300, share_MLP=True, max_words_l=16,
                         max_video_l=200, clip_len=0.17, max_ts_val=150.0,
                         wl=dict(name="charades_sta_vgg+glove_inference", baseline="configs[2]", oracle="charades_vgg", split=0,
                                 group_max=4, ragged=True, dur=(10.0, 60.0), dense_nms=0,
                                 algo_flops=2.72e9, algo_bytes=3297816)),
    "tacos": dict(_COMMON, dataset_name="tacos", v_feat_dim=4098, t_feat_dim=300, share_MLP=False, max_words_l=16,
                  max_video_l=200, clip_len=-1.0, max_ts_val=1000.0,
                  wl=dict(name="tacos_c3d_inference_dense_nms", baseline="configs[3]", oracle="tacos", split=0,
                          group_max=10, ragged=True, dur=(100.0, 900.0), dense_nms=100,
                          algo_flops=2.72e9, algo_bytes=3297816)),
    "qvh": dict(_COMMON, dataset_name="qvhighlights", v_feat_dim=2818, t_feat_dim=512, share_MLP=True, max_words_l=32,
                max_video_l=75, clip_len=2.0, max_ts_val=150.0,
                wl=dict(name="qvhighlights_c+sf_inference", baseline="configs[0] shape (CPU case B=32) / configs[4] shape", oracle="qvhighlights",
                        split=512, group_max=1, ragged=False, dur=(150.0, 150.0), dense_nms=0,
                        algo_flops=1.043e9, algo_bytes=911043)),
}
CHARADES_CSF = {k: v for k, v in BENCH_CONFIGS["charades_csf"].items() if k != "wl"}      # (tests import this)
NMS_THD = 0.7           # shipped configs disable NMS (nms_thd -1); 0.7 is Moment-DETR's convention


def model_cfg(name):
    return {k: v for k, v in BENCH_CONFIGS[name].items() if k != "wl"}


def make_workload(cfg, B, seed, device, wl=None):
    """Synthetic batch shaped like the collate output: video groups of 1..group_max queries that share their video,
    ragged Lv ~ U{Lv/2..Lv} (or uniform), per-source L2-normalised N(0,1) features + tef columns, words N(0,1) with lengths
    U{3..Lt}, durations U(dur), a synthetic ground-truth clip mask for the alignment scores."""
    wl = wl or BENCH_CONFIGS["charades_csf"]["wl"]
    g = torch.Generator(device="cpu").manual_seed(seed)
    Lv, Lt, Dv, Dt = cfg["max_video_l"], cfg["max_words_l"], cfg["v_feat_dim"], cfg["t_feat_dim"]
    nc = []
    while sum(nc) < B:
        nc.append(min(int(torch.randint(1, wl["group_max"] + 1, (1,), generator=g)), B - sum(nc)))
    G = len(nc)
    vlen_g = torch.randint(Lv // 2, Lv + 1, (G,), generator=g) if wl["ragged"] else torch.full((G,), Lv)
    vlen_g[0] = Lv
    nct = torch.tensor(nc)
    vlen = torch.repeat_interleave(vlen_g, nct)
    lo, hi = wl["dur"]
    dur = torch.repeat_interleave(torch.rand(G, generator=g) * (hi - lo) + lo, nct).float()
    wlen = torch.randint(3, Lt + 1, (B,), generator=g)
    gd = torch.Generator(device=device).manual_seed(seed)
    video = torch.empty(B, Lv, Dv, dtype=torch.float32, device=device)
    ar = torch.arange(Lv, device=device)
    split = wl["split"]
    start = 0
    for gi in range(0, G, 64):                      # build group-wise so that queries of a group share the video
        gs = list(range(gi, min(gi + 64, G)))
        x = torch.randn(len(gs), Lv, Dv - 2, device=device, generator=gd)
        if split:
            x[..., :split] = torch.nn.functional.normalize(x[..., :split], dim=-1)
            x[..., split:] = torch.nn.functional.normalize(x[..., split:], dim=-1)
        else:
            x = torch.nn.functional.normalize(x, dim=-1)
        L = vlen_g[gs].to(device).float()[:, None]
        tef = torch.stack([ar[None] / L, (ar[None] + 1) / L], dim=-1)
        x = torch.cat([x, tef], dim=-1) * (ar[None] < L)[..., None]
        rep = nct[gs].to(device)
        xr = torch.repeat_interleave(x, rep, dim=0)
        video[start:start + xr.shape[0]] = xr
        start += xr.shape[0]
    mask = ar[None] < vlen.to(device)[:, None]
    words = torch.randn(B, Lt, Dt, device=device, generator=gd)
    words = words * (torch.arange(Lt, device=device)[None] < wlen.to(device)[:, None])[..., None]
    clip_mask = (torch.rand(B, Lv, device=device, generator=gd) < 0.3) & mask      # GT clips of model/criterion.py:245 (synthetic)
    clip_mask[:, 0] = True
    out = dict(video_feat=video, video_mask=mask, words_feat=words, num_clips=nct, duration=dur.to(device), clip_mask=clip_mask,
               video_len=vlen.to(torch.int32))          # clip counts stay on the host (collate's `lengths`)
    if G >= 2:                                          # negative index: another video group (vectorised sample_outclass_neg)
        from mesm_b200.model import sample_outclass_neg
        out["neg_index"] = sample_outclass_neg(nct, generator=g).to(device)
    if wl["dense_nms"]:                                 # a 100-candidate / pair list for the NMS kernel
        n = wl["dense_nms"]
        st = torch.rand(B, n, device=device, generator=gd, dtype=torch.float64) * out["duration"][:, None].double() * 0.9
        w = torch.rand(B, n, device=device, generator=gd, dtype=torch.float64) * out["duration"][:, None].double() * 0.3
        sc = torch.round(torch.rand(B, n, device=device, generator=gd, dtype=torch.float64) * 1e4) / 1e4
        out["dense_windows"] = torch.stack([st, st + w, sc], dim=-1).reshape(B * n, 3).contiguous()
        out["dense_offsets"] = (torch.arange(B + 1, device=device, dtype=torch.int64) * n)
    return out


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  In-process NVML queries (no nvidia-smi process per sample: spawning
    one every 200 ms measurably slowed the step); falls back to nvidia-smi when pynvml is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows, self.stop, self.index, self.query_s = [], threading.Event(), index, []
        self.t = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        self.interval = float(os.environ.get("MESM_CLOCK_SAMPLE_S", "0.1"))     # every NVML query briefly stalls kernel submission
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            t0 = time.perf_counter()
            sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.query_s.append(time.perf_counter() - t0)
            return [str(sm), str(self.max_sm)] + ["Active" if mask & bit else "Not Active" for _, bit in self.REASONS]
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                           capture_output=True, text=True, timeout=5).stdout.strip()
        return [c.strip() for c in o.split(",")] if o else None

    def run(self):
        while not self.stop.is_set():
            try:
                r = self.sample()
                if r:
                    self.rows.append(r)
            except Exception:
                pass
            self.stop.wait(self.interval if self.nvml is not None else 0.5)

    def __enter__(self):
        if not os.environ.get("MESM_NO_CLOCK_SAMPLER"):
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.t.is_alive():
            self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        reasons = set()
        for r in self.rows:
            for (name, _), v in zip(self.REASONS, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]), "reasons": sorted(reasons), "samples": len(sm),
               "source": "nvml" if self.nvml is not None else "nvidia-smi"}
        if self.query_s:
            out["query_ms_mean"] = 1e3 * sum(self.query_s) / len(self.query_s)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# reference legs (checker code: the only places bench.py executes anything under oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_cfg(name, cfg):
    from oracle.config import CONFIGS
    return dataclasses.replace(CONFIGS[BENCH_CONFIGS[name]["wl"]["oracle"]], max_video_l=cfg["max_video_l"])


def _ref_align_scores(out, clip_mask, tau):
    """model/criterion.py:241-266 up to cos_sim / tau, on the reference's output dict."""
    import torch.nn.functional as F
    cm = clip_mask.unsqueeze(-1)
    cf = (out["projed_video_feat"] * cm).sum(1) / cm.sum(1)
    wm = out["expanded_words_mask"].unsqueeze(-1)
    wf = (out["expanded_words_feat"] * wm).sum(1) / wm.sum(1)
    return F.normalize(cf, dim=-1, p=2) @ F.normalize(wf, dim=-1, p=2).permute(1, 0) / tau


class ReferencePath:
    """The reference's own modules (oracle/_ref, staged by oracle/make_ref.py) driven as eval.py drives them - or, when they did
    not travel, the oracle port.  One call = forward (+ negative branch) + alignment scores + decode / post-process / NMS."""

    def __init__(self, name, cfg, state_dict, device="cpu"):
        from oracle import ref_harness
        self.name, self.cfg, self.ocfg, self.device = name, cfg, _oracle_cfg(name, cfg), torch.device(device)
        self.kind = "port"
        self.sd = {k: v.detach().float().to(self.device) for k, v in state_dict.items()}
        if ref_harness.reference_root() is not None and not os.environ.get("MESM_REF_PORT"):
            try:
                ref_harness.import_reference()
                self.model = ref_harness.build_reference(self.ocfg)
                self.model.load_state_dict(self.sd, strict=True)
                self.model = self.model.to(self.device).eval()
                self.kind = "reference"
            except Exception as e:                                   # noqa: BLE001 - any import problem -> the port
                print(f"[bench] reference modules unusable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        if self.kind == "port" and self.device.type != "cpu":
            raise RuntimeError("the oracle port runs on the CPU only")

    def batch(self, wl):
        d = self.device
        return {k: (v.to(d) if torch.is_tensor(v) and k not in ("num_clips", "video_len") else v) for k, v in wl.items()}

    def __call__(self, b, decode=True):
        from oracle import decode_oracle, mesm_oracle, ref_harness
        ocfg, dense = self.ocfg, self.cfg_wl()["dense_nms"]
        if self.kind == "reference":
            o = ref_harness.reference_forward(self.model, ocfg, b)
            S = _ref_align_scores(o, b["clip_mask"], ocfg.recss_tau)
            if decode:
                ref_harness.reference_decode(o["pred_logits"], o["pred_spans"], b["duration"], ocfg, NMS_THD)
                if dense:
                    from utils import temporal_nms
                    rows = b["dense_windows"].cpu().reshape(-1, dense, 3).tolist()
                    for r in rows:
                        temporal_nms(r, nms_thd=NMS_THD, max_after_nms=10)
        else:
            o = mesm_oracle.mesm_forward(self.sd, ocfg, b["video_feat"], b["video_mask"], b["words_feat"], b["num_clips"],
                                         neg_index=b["neg_index"])
            S = mesm_oracle.align_scores(o["projed_video_feat"], b["clip_mask"], o["expanded_words_feat"], o["expanded_words_mask"],
                                         ocfg.recss_tau)
            if decode:
                lg, sp = o["pred_logits"].numpy(), o["pred_spans"].numpy()
                for i in range(lg.shape[0]):
                    decode_oracle.decode_pair(lg[i], sp[i], float(b["duration"][i]), ocfg.clip_len, ocfg.max_ts_val, NMS_THD, 10, 10)
                if dense:
                    for r in b["dense_windows"].reshape(-1, dense, 3).tolist():
                        decode_oracle.temporal_nms(r, NMS_THD, 10)
        return o, S

    def cfg_wl(self):
        return BENCH_CONFIGS[self.name]["wl"]

    def describe(self):
        return ("the reference's own model.MESM + utils (span_cxw_to_xx, PostProcessorDETR, temporal_nms) driven as eval.py:63-116,476-485"
                if self.kind == "reference" else "oracle port of the reference forward + decode/NMS") + ", torch fp32"


def cpu_reference_pairs_per_s(name, cfg, state_dict, batch, steps, warmup, threads):
    torch.set_num_threads(threads)
    ref = ReferencePath(name, cfg, state_dict, "cpu")
    b = _f32_features(ref.batch(batch))
    B = b["video_feat"].shape[0]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ref(b)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return B * len(times) / sum(times), sum(times) / len(times), ref


def gpu_eager_pairs_per_s(name, cfg, state_dict, wl_full, pairs, device, steps=3):
    """The GPU bar of SURVEY 8d / BASELINE.md: the reference's modules in PyTorch eager mode on the same B200, fp32 with TF32 off."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref = ReferencePath(name, cfg, state_dict, device)
    b = _f32_features(ref.batch(take_groups(wl_full, pairs)))
    B = b["video_feat"].shape[0]
    res = {}
    for decode in (False, True):
        ref(b, decode=decode)                      # warm-up (lazy kernel loading, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            ref(b, decode=decode)
        torch.cuda.synchronize()
        res["forward_decode" if decode else "forward"] = B * steps / (time.perf_counter() - t0)
    return res, B, ref


def latency_b32(dev, feature_f16):
    """Small-batch latency (BASELINE configs[0] shape: QVHighlights C+SF, batch 32): one forward + decode/NMS, host call to results
    ready, eager (185+ launches enqueued from Python / ctypes) vs one CUDA-graph replay (Engine.capture)."""
    import mesm_b200
    from mesm_b200.model import build_model
    qcfg, qwl = model_cfg("qvh"), BENCH_CONFIGS["qvh"]["wl"]
    torch.manual_seed(0)
    m = build_model(qcfg).to(dev)
    b = make_workload(qcfg, 32, 99, dev, qwl)
    vf = b["video_feat"].half() if feature_f16 else b["video_feat"]
    vl = b["video_len"]

    def eager():
        o = m(vf, b["video_mask"], b["words_feat"], None, None, b["num_clips"], dataset_name="qvhighlights", is_training=False,
              neg_index=b["neg_index"], video_len=vl)
        return mesm_b200.decode_nms(o["pred_logits"], o["pred_spans"], b["duration"], qcfg["clip_len"], qcfg["max_ts_val"], NMS_THD, 10, 10)

    for _ in range(5):
        eager()
    torch.cuda.synchronize()
    lat = []
    for _ in range(30):
        t0 = time.perf_counter()
        w = eager()
        w[0][0, 0, 0].item()                       # results on the host
        lat.append(time.perf_counter() - t0)
    lat.sort()
    eng = m._eng
    cap = eng.capture(vf, b["video_mask"], b["words_feat"], b["num_clips"], neg_index=b["neg_index"], want=("core", "rec", "aux"), video_len=vl,
                      decode=dict(duration=b["duration"], clip_len=qcfg["clip_len"], max_ts_val=qcfg["max_ts_val"], nms_thd=NMS_THD))
    ref = eager()
    out = cap.replay()
    torch.cuda.synchronize()
    same = bool(torch.equal(out["windows"], ref[0]) and torch.equal(out["keep"], ref[2]))
    glat = []
    for _ in range(50):
        t0 = time.perf_counter()
        o = cap.replay()
        o["windows"][0, 0, 0].item()
        glat.append(time.perf_counter() - t0)
    glat.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        cap.replay()
    e1.record()
    torch.cuda.synchronize()
    return {"shape": "QVHighlights C+SF, batch 32, forward (incl. negative branch) + decode/NMS, host call -> first result on the host",
            "eager_ms_median": 1e3 * lat[len(lat) // 2], "eager_ms_min": 1e3 * lat[0],
            "graph_ms_median": 1e3 * glat[len(glat) // 2], "graph_ms_min": 1e3 * glat[0],
            "graph_device_ms_back_to_back": e0.elapsed_time(e1) / 50, "graph_kernels": cap.launches, "graph_equals_eager": same}


def _f32_features(b):
    """The reference legs take fp32 tensors: the stored fp16 values upcast (exactly the values our arm computes on)."""
    for k in ("video_feat", "words_feat"):
        if b[k].dtype == torch.float16:
            b = dict(b, **{k: b[k].float()})
    return b


def take_groups(wl, n_pairs):
    """First whole video groups covering >= n_pairs pairs (sample of the same workload)."""
    nc = wl["num_clips"].tolist()
    tot, k = 0, 0
    while tot < n_pairs and k < len(nc):
        tot += nc[k]
        k += 1
    if k < 2 and len(nc) >= 2:                     # the negative branch needs two groups
        k = 2
        tot = nc[0] + nc[1]
    from mesm_b200.model import sample_outclass_neg
    sub = {key: wl[key][:tot] for key in ("video_feat", "video_mask", "words_feat", "duration", "video_len", "clip_mask")}
    sub["num_clips"] = wl["num_clips"][:k]
    sub["neg_index"] = sample_outclass_neg(sub["num_clips"], generator=torch.Generator().manual_seed(5)).to(wl["video_feat"].device)
    if "dense_windows" in wl:
        n = wl["dense_windows"].shape[0] // wl["video_feat"].shape[0]
        sub["dense_windows"] = wl["dense_windows"][:tot * n]
        sub["dense_offsets"] = wl["dense_offsets"][:tot + 1]
    return sub


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="charades_csf", choices=sorted(BENCH_CONFIGS))
    ap.add_argument("--pairs", type=int, default=4096, help="pairs per GPU per step")
    ap.add_argument("--chunk-pairs", type=int, default=0, help="pairs per internal chunk (0 = engine default)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=32)
    ap.add_argument("--eager-pairs", type=int, default=1024, help="pairs of the PyTorch-eager-on-GPU reference leg (0 = skip)")
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--e2e-sub", type=int, default=2048, help="pairs per host->device sub-batch of the e2e measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--feature-dtype", default="f16", choices=["f16", "f32"],
                    help="storage format of the clip features (host and HBM): f16 = the 16-bit storage option of the ingest front-end "
                         "(values used exactly; every arm - ours, the CPU reference, the eager GPU reference - is fed the same values), "
                         "f32 = fp32 tensors as the reference's loaders deliver them")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cfg, wlc = model_cfg(args.config), BENCH_CONFIGS[args.config]["wl"]
    config = {"workload": wlc["name"] if args.pairs == 4096 or args.config != "charades_csf" else f"charades_sta_c+sf_inference_batch{args.pairs}",
              "config": wlc["baseline"], "pairs_per_gpu": args.pairs,
              "Lv": cfg["max_video_l"], "Lt": cfg["max_words_l"], "v_feat_dim": cfg["v_feat_dim"], "t_feat_dim": cfg["t_feat_dim"],
              "ragged_video": f"U{{{cfg['max_video_l'] // 2}..{cfg['max_video_l']}}}" if wlc["ragged"] else "uniform",
              "negative_branch": True, "align_scores": True, "nms_thd": NMS_THD, "parallelism": f"dp{world}",
              "feature_storage": "fp16 (fp16-representable clip and word features, used exactly; computed in fp32 accumulate)" if args.feature_dtype == "f16" else "fp32",
              "l2": "inputs (GBs per GPU) larger than L2; no flush needed"}
    if wlc["dense_nms"]:
        config["dense_nms_candidates"] = wlc["dense_nms"]

    from mesm_b200.model import build_model
    torch.manual_seed(0)
    model = build_model(cfg)                       # random-init weights of the architecture (no checkpoints offline)
    with torch.no_grad():                          # exercise the tensors the reference initialises to constants
        for n, p in model.named_parameters():
            if n.endswith("bbox_embed.layers.2.weight") or n.endswith("masked_sent_token"):
                p.normal_(0, 0.02)

    # ------------------------------------------------------------------ reference arm: host CPU cores -----------------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        wl = make_workload(cfg, 64, 1234, "cpu", wlc)
        if args.feature_dtype == "f16":
            wl["video_feat"] = wl["video_feat"].half().float()       # the same fp16-representable values our arm stores in 16 bits
            wl["words_feat"] = wl["words_feat"].half().float()
        sub = take_groups(wl, args.cpu_sample_pairs)
        pps, sec, ref = cpu_reference_pairs_per_s(args.config, cfg, model.state_dict(), sub, max(args.steps, 1), max(args.warmup, 0), threads)
        line = {"impl": "reference", "metric": "video-query pairs/sec", "value": pps, "unit": "pairs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": pps, "unit": "pairs/s", "cores": threads, "kind": ref.kind,
                                 "sample": f"{sub['video_feat'].shape[0]} pairs of the same workload per step ({ref.describe()}, all host threads)"},
                "e2e": {"value": pps, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm ----------------------------------------
    import mesm_b200
    from mesm_b200 import _lib
    # host placement before any pinned allocation: this rank's threads and staging buffers on its GPU's NUMA node
    numa = {"bound": False, "why": "MESM_NO_NUMA_BIND"} if os.environ.get("MESM_NO_NUMA_BIND") else mesm_b200.bind_to_gpu_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    model = model.to(dev)
    model.chunk_pairs = args.chunk_pairs
    wl = make_workload(cfg, args.pairs, 1234 + rank, dev, wlc)
    f16 = args.feature_dtype == "f16"
    if f16:
        wl["video_feat"] = wl["video_feat"].half()                   # resident in HBM in the storage format
        wl["words_feat"] = wl["words_feat"].half()                   # (the engine widens the 16-bit word features on the device)
    B, Lv = args.pairs, cfg["max_video_l"]
    lib = _lib.lib()
    from mesm_b200.sharding import gather_topk

    # host clip counts (what the collate step knows): the engine then runs on packed variable-length rows
    vlen_host = None if os.environ.get("MESM_PADDED_ROWS") else wl["video_len"]
    layout = {"rows": "zero-padded [B, Lv]" if vlen_host is None else "packed variable-length (host clip counts passed as video_len)",
              "clip_rows_per_gpu": int(B * Lv if vlen_host is None else wl["video_len"].sum())}
    dname = cfg["dataset_name"]
    # the collate step replicates a group's video for each of its queries (dataset/base.py:307-309): the engine may read a pair's
    # clips from its group's first pair and project every video once (charades / tacos grouping only)
    shared_res = vlen_host is not None and dname != "qvhighlights" and not os.environ.get("MESM_NO_SHARED")
    layout["video_projection"] = "once per video (shared_group_video)" if shared_res else "once per pair"

    def step():
        out = model(wl["video_feat"], wl["video_mask"], wl["words_feat"], None, None, wl["num_clips"],
                    dataset_name=dname, is_training=False, neg_index=wl["neg_index"], video_len=vlen_host,
                    shared_group_video=shared_res)
        S = mesm_b200.align_scores(out["projed_video_feat"], wl["clip_mask"], out["expanded_words_feat"], out["expanded_words_mask"], 0.5)
        win, order, keep, cnt = mesm_b200.decode_nms(out["pred_logits"], out["pred_spans"], wl["duration"], cfg["clip_len"],
                                                     cfg["max_ts_val"], NMS_THD, 10, 10)
        n = 3
        if wlc["dense_nms"]:
            mesm_b200.temporal_nms_lists(wl["dense_windows"], wl["dense_offsets"], NMS_THD, 10)
            n += 1
        return out, win, order, keep, cnt, S, n

    _v = lambda m: print(f"[bench] {m}", file=sys.stderr, flush=True) if os.environ.get("MESM_BENCH_VERBOSE") else None
    for _ in range(args.warmup):
        o_w = step()
    gather_topk(o_w[1], wl["num_clips"], args.topk, rank, world, B)     # first use of the sort kernels / NCCL communicator
    del o_w
    torch.cuda.synchronize()
    _v("warmup done")
    # Everything built so far (modules, workload, ctypes tables) is long-lived: park it in the permanent generation so that a
    # full collection triggered by the per-step garbage (lists of clip counts, ctypes arrays) cannot stall the host for tens
    # of milliseconds in the middle of a timed region (seen as one 55 ms forward() in MESM_E2E_TRACE runs).
    gc.collect()
    gc.freeze()
    if dist:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        ev0.record()
        step_ev = [ev0]
        for _ in range(args.steps):
            out, win, order, keep, cnt, S, n_extra = step()
            launches += model._eng.last_launch_count + n_extra
            step_ev.append(torch.cuda.Event(enable_timing=True))
            step_ev[-1].record()
        top = gather_topk(win, wl["num_clips"], args.topk, rank, world, B)      # one NCCL all_gather of top-k spans
        ev1.record()
        torch.cuda.synchronize()
    if dist:
        dist.barrier()
    _v("timed region done")
    ms = ev0.elapsed_time(ev1)
    each_ms = [round(a.elapsed_time(b), 2) for a, b in zip(step_ev[:-1], step_ev[1:])]
    t = torch.tensor([ms], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = world * B * args.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel (fused FFN): one extra step with CUDA events around each launch ----------
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.mesm_profile_begin()
    pe0.record()
    step()
    pe1.record()
    prof = (ctypes.c_double * 7)()
    lib.mesm_profile_end(prof)
    torch.cuda.synchronize()
    profile_step_ms = pe0.elapsed_time(pe1)
    _v("profile step done")
    rep_rows = {r[0]: (int(r[1]), float(r[2])) for r in (l.split("\t") for l in lib.mesm_profile_report().decode().strip().split("\n")) if len(r) >= 3}
    if os.environ.get("MESM_PROFILE_REPORT"):
        for k, (n, t_ms) in sorted(rep_rows.items(), key=lambda kv: -kv[1][1]):
            print(f"{t_ms:9.3f} ms  n={n:5d}  {k}", file=sys.stderr)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    lin_ms, lin_flops = prof[0], prof[1]
    step_ms = ms / args.steps
    clip_rows = float(layout["clip_rows_per_gpu"])      # rows the kernels actually process
    traffic_tab = {}
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                traffic_tab = json.load(f)
            break
        except Exception:
            pass
    # dominant kernel = the fused FFN (ffn_pair_kernel): 8 T2V launches on the clip rows + 4 encoder launches on clip rows + B per step
    ffn_n, ffn_ms = rep_rows.get("ffn_fused", (0, 0.0))
    if ffn_n:
        ffn_flops = 4.0 * 256 * 1024 * (8.0 * clip_rows + 4.0 * (clip_rows + B))      # 12 layer passes over the step's rows (however chunked)
        ach = ffn_flops / (ffn_ms * 1e-3) / 1e12
        tr = traffic_tab.get("ffn_pair_kernel", {}).get("dram_bytes_per_row")
        roof = {"bound": "tensor", "kernel": "ffn_pair_kernel (fused FFN block, tcgen05 CTA pairs; bf16x3 = 3 MMAs per algorithmic MAC)",
                "achieved": ach, "peak": peak_tf, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback B200_PROFILING.md",
                "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": tr * clip_rows if tr else None, "launches_per_step": ffn_n,
                "kernel_ms_per_step": ffn_ms, "share_of_step": ffn_ms / step_ms, "issued_frac": 3 * ach / peak_tf}
    else:
        roof = {"bound": "tensor", "kernel": "tcgen05 linear (all GEMM launches of one step)", "achieved": None, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": None, "traffic": None}
    roof["all_gemm_launches"] = {"achieved": lin_flops / (lin_ms * 1e-3) / 1e12 if lin_ms else None,
                                 "frac": (lin_flops / (lin_ms * 1e-3) / 1e12 / peak_tf) if lin_ms else None,
                                 "launches_per_step": int(prof[3]), "kernel_ms_per_step": lin_ms, "share_of_step": lin_ms / step_ms}
    # second entry: the one HBM-bound stage - the K = Dv input projection that streams the clip features
    k1 = [(k, v) for k, v in rep_rows.items() if k.startswith("input_proj")]
    if k1:
        k1_ms = sum(v[1] for _, v in k1)
        feat_bytes = float(model._eng.last_feature_bytes)
        roof["input_projection"] = {"bound": "hbm", "kernel": k1[0][0], "achieved": feat_bytes / (k1_ms * 1e-3) / 1e9, "peak": peak_hbm,
                                    "unit": "GB/s", "frac": feat_bytes / (k1_ms * 1e-3) / 1e9 / peak_hbm, "kernel_ms_per_step": k1_ms,
                                    "algorithmic_bytes": feat_bytes}
    roof["profiled_step_ms"] = profile_step_ms
    roof["profiled_kernels_ms"] = sum(v[1] for v in rep_rows.values())
    roof["hbm_frac_whole_step"] = (wlc["algo_bytes"] * B / (step_ms * 1e-3)) / 1e9 / peak_hbm

    # ---- e2e: same work through the public API from pinned host memory, sub-batches double-buffered over two streams -----
    sub = args.e2e_sub if B >= args.e2e_sub else B
    nsub = max(1, B // sub)
    sb = take_groups(wl, sub)                      # sub-batches hold whole groups: a slice with its own grouping
    Bs = sb["video_feat"].shape[0]
    host_keys = ["video_feat", "video_mask", "words_feat", "duration", "neg_index", "clip_mask"] + (["dense_windows", "dense_offsets"] if wlc["dense_nms"] else [])
    host = {k: sb[k].cpu().pin_memory() for k in host_keys}
    padded = bool(os.environ.get("MESM_E2E_PADDED"))            # A/B switch: plain copy of the zero-padded tensor
    # One context = one compute stream (a mesm_ctx is not re-entrant).  A second stream prefetches the next sub-batch from
    # pinned host memory while the current one is being scored; events order copy -> compute -> buffer reuse.  Steps are
    # streamed back to back (the first sub-batch of step k+1 is prefetched under the last sub-batch of step k).
    # The host->device step is the drop-in of the reference's prepare_batch_input (dataset/base.py:358): only the valid
    # clip rows of each pair cross PCIe, the pad rows are zero-filled on the device (mesm_upload_clips).
    # (the copy stream has the higher priority: its one small kernel - the pad-row zero fill - must not queue behind the
    # compute stream's grids, or the copy engine idles until it has run)
    comp, copy = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    shared = (not padded and vlen_host is not None and not os.environ.get("MESM_E2E_NO_SHARED") and dname != "qvhighlights")
    # NB device buffers: with two, the copy of sub-batch i+2 can only start when the scoring of sub-batch i has finished, so the copy
    # engine idles whenever a copy is shorter than a scoring period and stalls the GPU whenever it is longer; a third buffer lets the
    # copy stream run continuously (at N = 8 the host bridges of this box leave some GPUs only ~23 GB/s, profiles/r2_h2d_probe_n8.md)
    NB = int(os.environ.get("MESM_E2E_BUFFERS", "3"))
    dbuf = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(NB)]
    hres = [torch.empty(Bs, 10, 3, dtype=torch.float64).pin_memory() for _ in range(NB)]
    hkeep = [torch.empty(Bs, 10, dtype=torch.int32).pin_memory() for _ in range(NB)]
    ready = [torch.cuda.Event() for _ in range(NB)]
    freed = [torch.cuda.Event() for _ in range(NB)]
    h2d_box = [0]
    vl_box = [None if vlen_host is None else sb["video_len"]] * NB

    # ---- NVLink-assisted ingest (mesm_b200/relay.py): ranks on a slow host link route part of their videos through a fast peer ----
    relay, relay_info = None, None
    if dist and shared and not padded and not os.environ.get("MESM_NO_RELAY"):
        try:
            ff = os.environ.get("MESM_RELAY_FORCE")
            relay, relay_info = mesm_b200.plan_ingest_relay(local, dist, int(0.12 * host["video_feat"].numel() * host["video_feat"].element_size()),
                                                            nbuf=NB, force_fraction=float(ff) if ff else None)
        except Exception as e:                                   # noqa: BLE001 - the direct path always works
            relay, relay_info = None, {"error": f"{type(e).__name__}: {e}"}
    use_relay = [relay is not None]

    def e2e_stream(total):
        # Host order: scoring of sub-batch i is enqueued first, then the ingest of sub-batch i+2 into the buffer it frees.
        for e in freed:
            e.record(comp)

        def prefetch(i):
            with torch.cuda.stream(copy):
                copy.wait_event(freed[i % NB])                  # the compute that last read this buffer has finished
                if padded:
                    for k, v in host.items():
                        dbuf[i % NB][k].copy_(v, non_blocking=True)
                    h2d_box[0] = sum(v.numel() * v.element_size() for v in host.values())
                else:
                    staged = mesm_b200.prepare_batch_input(dict(host, num_clips=sb["num_clips"]), dev, non_blocking=True, out=dbuf[i % NB],
                                                           shared_group_video=shared, relay=relay if use_relay[0] else None)
                    h2d_box[0] = mesm_b200.prepare_batch_input.last_h2d_bytes
                    vl_box[i % NB] = None if vlen_host is None else staged["video_len"]
                ready[i % NB].record(copy)

        trace = [] if os.environ.get("MESM_E2E_TRACE") else None
        t_host0 = time.perf_counter()
        for j in range(min(NB, total)):
            prefetch(j)
        th = th1 = time.perf_counter()
        for i in range(total):
            d = dbuf[i % NB]
            tf0 = time.perf_counter()
            with torch.cuda.stream(comp):
                comp.wait_event(ready[i % NB])
                o = model(d["video_feat"], d["video_mask"], d["words_feat"], None, None, sb["num_clips"],
                          dataset_name=dname, is_training=False, neg_index=d["neg_index"], video_len=vl_box[i % NB],
                          shared_group_video=shared)
                mesm_b200.align_scores(o["projed_video_feat"], d["clip_mask"], o["expanded_words_feat"], o["expanded_words_mask"], 0.5)
                w, od, kp, ct = mesm_b200.decode_nms(o["pred_logits"], o["pred_spans"], d["duration"], cfg["clip_len"],
                                                     cfg["max_ts_val"], NMS_THD, 10, 10)
                if wlc["dense_nms"]:
                    mesm_b200.temporal_nms_lists(d["dense_windows"], d["dense_offsets"], NMS_THD, 10)
                hres[i % NB].copy_(w, non_blocking=True)
                hkeep[i % NB].copy_(kp, non_blocking=True)
                freed[i % NB].record(comp)
                if trace is not None:
                    e = torch.cuda.Event(enable_timing=True); e.record(comp)
                    trace.append((e, tf0 - t_host0, th1 - th, time.perf_counter() - tf0, model._eng.last_enqueue_s))
            th = time.perf_counter()
            if i + NB < total:
                prefetch(i + NB)
            th1 = time.perf_counter()
        comp.synchronize()
        copy.synchronize()
        if trace:
            for j in range(1, len(trace)):
                print(f"[e2e] sub {j}: gpu period {trace[j - 1][0].elapsed_time(trace[j][0]):7.2f} ms  host: t={trace[j][1] * 1e3:7.1f} "
                      f"prefetch {trace[j][2] * 1e3:6.2f} ms forward {trace[j][3] * 1e3:6.2f} ms (engine set-up {trace[j][4][0] * 1e3:.2f}, "
                      f"launch enqueue {trace[j][4][1] * 1e3:.2f})", file=sys.stderr)

    d2h = hres[0].numel() * 8 + hkeep[0].numel() * 4
    e2e_stream(max(2 * nsub, NB))    # warm-up: every buffer, the allocator's steady state and the copy path
    torch.cuda.synchronize()
    if relay_info is not None and relay_info.get("pairs"):
        # keep the relay only if it is faster for the NODE (max over ranks), measured on a few streamed steps each way
        def timed(on):
            use_relay[0] = on and relay is not None
            e2e_stream(NB)
            torch.cuda.synchronize()
            dist.barrier()
            t_ = time.perf_counter()
            e2e_stream(4 * nsub)
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t_], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt)
        def timed_mode(batched):
            if relay is not None:
                relay.batched = batched
            return timed(True)
        t_off, t_b, t_c = timed(False), timed_mode(True), timed_mode(False)
        relay_info["trial_s"] = {"direct": round(t_off, 4), "relayed_batched_copies": round(t_b, 4), "relayed_single_copies": round(t_c, 4)}
        t_on = min(t_b, t_c)
        relay_info["used"] = bool(t_on < 0.98 * t_off)
        relay_info["mode"] = "batched" if t_b <= t_c else "single"
        if relay is not None:
            relay.batched = t_b <= t_c
        use_relay[0] = relay_info["used"] and relay is not None
    elif relay_info is not None:
        relay_info["used"] = False
    gc.collect()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps)
    e2e_stream(nsub * e2e_steps)
    torch.cuda.synchronize()
    _v("e2e done")
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * Bs * nsub * e2e_steps / float(t)

    line = {"metric": "video-query pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_each_step": each_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "layout": layout, "numa": numa, "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d_box[0] * nsub, "d2h_bytes_per_step": d2h * nsub,
                    "note": f"{nsub} sub-batches of {Bs} pairs per step, {e2e_steps} steps streamed back to back; pinned host -> device ingest up to {NB} sub-batches ahead ({NB} device buffers) on a copy stream through mesm_b200.prepare_batch_input ({'zero-padded tensor copied whole' if padded else 'valid clip rows only, pad rows zero-filled on the device' + ('; the video a group of queries shares (replicated by the collate step, dataset/base.py:307-309) crosses PCIe once' if shared else '')}), windows + keep sets back to host", "h2d_padded_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()) * nsub},
            "gpu_launches": int(launches), "roofline": roof, "topk_gathered": int(top.shape[0])}
    if relay_info is not None:
        if relay is not None:
            relay_info["this_rank"] = relay.describe()
        line["e2e"]["ingest_relay"] = relay_info
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        cs = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in take_groups(wl, args.cpu_sample_pairs).items()}
        pps, sec, ref = cpu_reference_pairs_per_s(args.config, cfg, sd_cpu, cs, 3, 1, threads)
        line["cpu_baseline"] = {"value": pps, "unit": "pairs/s", "cores": threads, "kind": ref.kind,
                                "sample": f"{cs['video_feat'].shape[0]} pairs of the same workload, 3 timed iterations after 1 warm-up "
                                          f"({ref.describe()})"}
        del ref
        # SURVEY 8d C1 (the reference's own CPU-runnable case, BASELINE configs[0]): QVHighlights shape, batch 32
        try:
            qcfg, qwl = model_cfg("qvh"), BENCH_CONFIGS["qvh"]["wl"]
            torch.manual_seed(0)
            qsd = build_model(qcfg).state_dict()
            qb = make_workload(qcfg, 32, 77, "cpu", qwl)
            qpps, _, qref = cpu_reference_pairs_per_s("qvh", qcfg, qsd, qb, 3, 1, threads)
            line["cpu_baseline"]["c1_qvh_b32"] = {"value": qpps, "unit": "pairs/s", "kind": qref.kind,
                                                  "sample": "BASELINE configs[0]: QVHighlights C+SF shape, batch 32, forward + decode/NMS, 3 iterations"}
            del qref, qsd, qb
        except Exception as e:                      # noqa: BLE001
            line["cpu_baseline"]["c1_qvh_b32"] = {"error": f"{type(e).__name__}: {e}"}
        try:
            line["latency_b32"] = latency_b32(dev, f16)
        except Exception as e:                      # noqa: BLE001
            line["latency_b32"] = {"error": f"{type(e).__name__}: {e}"}
        if args.eager_pairs > 0:
            try:
                res, Be, gref = gpu_eager_pairs_per_s(args.config, cfg, sd_cpu, wl, min(args.eager_pairs, B), dev)
                line["gpu_eager_baseline"] = {"value": res["forward_decode"], "forward_only": res["forward"], "unit": "pairs/s", "kind": gref.kind,
                                              "sample": f"{Be} pairs of the same workload on the same B200: {gref.describe()}, PyTorch eager, TF32 off; "
                                                        "forward_only excludes the Python decode loop"}
                line["vs_gpu_eager"] = {"value_over_eager_forward_only": value / res["forward"], "value_over_eager_forward_decode": value / res["forward_decode"]}
            except Exception as e:                  # noqa: BLE001
                line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
