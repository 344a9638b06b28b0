"""CPU: C-ABI library loads and exports every declared symbol; host-side logic (state_dict layout, sharding, negative
sampling); world_size-2 gloo run of the top-k gather."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mesm_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mesm_b200.h")).read()
    declared = set(re.findall(r"\b(mesm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    l = _lib.lib()                       # dlopen + getattr of every symbol
    assert l.mesm_abi_version() == 3


def test_no_device_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mesm_b200
    with pytest.raises(RuntimeError):
        mesm_b200.Engine(dict(v_feat_dim=130, t_feat_dim=64))
    with pytest.raises(RuntimeError):
        mesm_b200.decode_nms(torch.zeros(1, 10, 2), torch.zeros(1, 10, 2), torch.ones(1), 2.0, 150.0)


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "mesm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle/|oracle\.", txt, re.M), f


@pytest.mark.parametrize("cfg_name", ["qvhighlights", "charades_csf", "charades_vgg", "tacos"])
def test_state_dict_keys_match_reference(cfg_name):
    from oracle.config import CONFIGS
    from oracle.weights import state_spec
    from mesm_b200.model import build_model
    c = CONFIGS[cfg_name]
    d = c.asdict()
    d["share_MLP"] = d.pop("share_mlp")
    m = build_model(d)
    spec = dict(state_spec(c))           # pinned against the real reference module by oracle/gen_golden.py
    sd = m.state_dict()
    assert set(sd) == set(spec)
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)


def test_shard_groups_partition():
    from mesm_b200.sharding import shard_batch, shard_groups
    nc = [2, 1, 3, 1, 2, 4, 1, 2, 5, 1]
    B = sum(nc)
    for world in (1, 2, 4, 8):
        covered, prev = [], 0
        for r in range(world):
            glo, ghi, plo, phi = shard_groups(nc, r, world)
            assert plo == prev and phi - plo == sum(nc[glo:ghi])
            prev = phi
            covered += list(range(plo, phi))
        assert covered == list(range(B))
    batch = dict(num_clips=torch.tensor(nc), x=torch.arange(B), y=torch.arange(B * 2).view(B, 2))
    parts = [shard_batch(batch, r, 4) for r in range(4)]
    assert torch.equal(torch.cat([p["x"] for p in parts]), batch["x"])
    assert torch.equal(torch.cat([p["num_clips"] for p in parts]), batch["num_clips"])


def test_topk_merge_equals_single_rank():
    from mesm_b200.sharding import local_topk, merge_topk
    g = torch.Generator().manual_seed(0)
    B, nq, k = 64, 10, 7
    win = torch.rand(B, nq, 3, generator=g, dtype=torch.float64)
    win[:, :, 2] = (win[:, :, 2] * 50).round() / 50                     # many score ties
    order = torch.randint(0, nq, (B, nq), generator=g, dtype=torch.int32)
    full = local_topk(win, order, k)
    for world in (2, 4, 8):
        per = B // world
        recs = torch.cat([local_topk(win[r * per:(r + 1) * per], order[r * per:(r + 1) * per], k, r * per) for r in range(world)])
        assert torch.equal(merge_topk(recs, k), full)


def test_sample_outclass_neg_other_group():
    from mesm_b200.model import sample_outclass_neg
    nc = torch.tensor([2, 1, 3, 1, 4])
    grp = torch.repeat_interleave(torch.arange(len(nc)), nc)
    for s in range(20):
        neg = sample_outclass_neg(nc, generator=torch.Generator().manual_seed(s))
        assert neg.min() >= 0 and neg.max() < int(nc.sum()) and bool((grp[neg] != grp).all())
    with pytest.raises(IndexError):
        sample_outclass_neg(torch.tensor([3]))


def test_gloo_world2_topk_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from mesm_b200.sharding import gather_topk, local_topk
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
g = torch.Generator().manual_seed(0)
win = torch.rand(32, 10, 3, generator=g, dtype=torch.float64)
per = 32 // w
top = gather_topk(win[r*per:(r+1)*per], None, 5, r, w, per)
full = local_topk(win, torch.zeros(32, 10, dtype=torch.int32), 5)
assert torch.equal(top, full), (top, full)
dist.destroy_process_group()
print('ok', r)
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert p.returncode == 0, p.stdout + p.stderr


def test_clip_counts_are_the_collate_lengths():
    """ingest.clip_counts: index of the last valid clip + 1 (= `lengths` of utils/data_utils.py:64 for prefix masks), >= 1."""
    from mesm_b200.ingest import clip_counts
    L = 9
    lens = torch.tensor([9, 1, 4, 7])
    mask = torch.arange(L)[None] < lens[:, None]
    assert clip_counts(mask).tolist() == lens.tolist() and clip_counts(mask).dtype == torch.int32
    holes = torch.tensor([[1, 1, 0, 1, 0, 0], [0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 0, 1]], dtype=torch.bool)
    assert clip_counts(holes).tolist() == [4, 1, 6, 6]
    assert clip_counts(mask.float()).tolist() == lens.tolist()          # the collate's float mask before .bool()


def test_header_and_binding_agree_on_struct_layouts():
    """mesm_inputs / mesm_cfg field order in include/mesm_b200.h == the ctypes Structures (a silent mismatch would shift pointers)."""
    from mesm_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mesm_b200.h")).read()
    def fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return [re.search(r"(\w+)\s*;", ln).group(1) for ln in body.split("\n") if ";" in ln]
    assert fields("mesm_inputs") == [f[0] for f in _lib.MesmInputs._fields_]
    assert fields("mesm_cfg") == [f[0] for f in _lib.MesmCfg._fields_]
    assert fields("mesm_outputs") == [f[0] for f in _lib.MesmOutputs._fields_]
    assert fields("mesm_decode_params") == [f[0] for f in _lib.MesmDecodeParams._fields_]


def test_relay_pairing_of_asymmetric_host_links():
    """mesm_b200.relay.pair_links: slowest rank with fastest, share that equalises the two links; symmetric nodes get no relay."""
    from mesm_b200.relay import pair_links
    assert pair_links([55.5, 55.6]) == []
    assert pair_links([36.0, 35.0, 36.5, 35.5]) == []
    pairs = pair_links([23.3, 23.3, 28.8, 28.5, 36.2, 36.3, 38.6, 39.7])
    assert [(p["rank"], p["via_rank"]) for p in pairs] == [(0, 7), (1, 6), (3, 5), (2, 4)]
    for p, (s, f) in zip(pairs, [(23.3, 39.7), (23.3, 38.6), (28.5, 36.3), (28.8, 36.2)]):
        x = p["fraction"]
        assert abs((1 - x) / s - (1 + x) / f) < 2e-4          # both links finish together
    assert all(p["fraction"] <= 0.3 for p in pair_links([10.0, 40.0]))
    assert pair_links([50.0, 50.0], force_fraction=0.2) == [{"rank": 0, "via_rank": 1, "fraction": 0.2}]
