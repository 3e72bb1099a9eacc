"""dalle2_video_b200 — B200-native decoder denoising step for DALLE2-video.

Public surface (mirrors /root/reference/dalle2_video/dalle2_video.py):
    Unet3D, VideoDecoder, NoiseScheduler, NullVQGanVAE
plus `shard_rows` for batch-sharded multi-GPU sampling (no collective on the data path) and `AttentionBlock`
(stand-alone joint / factorised / temporal attention blocks for the configs[4] kernel sweep).
`ViViT` mirrors dalle2_video/video_encoder.py (forward only: the producer of `video_embed`).
The CUDA extension (libdv_b200.so, built by __graft_entry__.build()) is mandatory.
"""
from .attention import AttentionBlock
from .decoder import NullVQGanVAE, VideoDecoder
from .parallel import shard_rows
from .scheduler import NoiseScheduler
from .train import VideoDecoderTrainer
from .unet3d import Unet3D
from .vivit import ViViT

__all__ = ["Unet3D", "VideoDecoder", "NoiseScheduler", "NullVQGanVAE", "shard_rows", "AttentionBlock", "VideoDecoderTrainer", "ViViT"]
