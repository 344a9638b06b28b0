"""mesm_b200 — B200-native (sm_100a) implementation of MESM's per-pair cross-modal inference path.

Layout: ``csrc/`` CUDA kernels + C ABI (include/mesm_b200.h) -> ``libmesm_b200.so``; ``_lib`` ctypes binding;
``engine`` context owner; ``model`` / ``utils`` drop-in mirrors of the reference's Python interface.
"""
from .engine import Engine, decode_nms, temporal_nms_lists, align_scores, loss_saliency  # noqa: F401
from .ingest import prepare_batch_input, upload_clips, build_video_feat  # noqa: F401
from .numa import bind_to_gpu_node  # noqa: F401
from .relay import IngestRelay, plan_ingest_relay  # noqa: F401

__all__ = ["Engine", "decode_nms", "temporal_nms_lists", "align_scores", "loss_saliency", "prepare_batch_input", "upload_clips",
           "build_video_feat", "bind_to_gpu_node", "IngestRelay", "plan_ingest_relay"]
