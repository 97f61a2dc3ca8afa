"""edsnet_b200 -- B200 (sm_100a) implementation of EDSNet's anchor-based scoring path.

Public surface:
  DSNet            drop-in for the reference's `anchor_based.dsnet.DSNet` (nystromformer + roi pooling)
  BatchPlan        packed variable-length batch tables
  shard_videos     video-wise partition across ranks (no collective on the data path)
  ScoringPipeline  host-buffer -> proposals throughput path (pinned H2D / compute / D2H overlapped)
  ShotPlan / keyshot_summaries   kept proposals -> keyshot summaries on the device (bbox2summary)
  TruthPlan / eval_metrics / evaluate   F-score and diversity of the summaries on the device (evaluate.py)
  kts_change_points / kts_shots   kernel temporal segmentation (shot boundaries) on the device
  GoogLeNetPool5   pool5 frame features (the reference's FeatureExtractor('google-net')) on the tcgen05 GEMM
  summarize        infer.py's chain from the sampled features on: segmentation -> scores -> NMS -> keyshot summary
  training         anchor labels, cls/loc losses, data-parallel step with one flat gradient all-reduce
"""
from .plan import BatchPlan, DeviceBatch, shard_videos          # noqa: F401
from .dsnet import DSNet, NystromAttention, AttentionExtractor                      # noqa: F401
from .pipeline import ScoringPipeline                            # noqa: F401
from . import training                                           # noqa: F401
from .summary import ShotPlan, keyshot_summaries, keyshot_from_scores, training_targets, split_summaries  # noqa: F401
from .evaluate import TruthPlan, eval_metrics, evaluate            # noqa: F401
from .kts import kts_change_points, kts_shots                      # noqa: F401
from .infer import summarize, summarize_frames                                       # noqa: F401
from .features import GoogLeNetPool5                               # noqa: F401

__all__ = ["DSNet", "NystromAttention", "BatchPlan", "DeviceBatch", "shard_videos", "ScoringPipeline"]
