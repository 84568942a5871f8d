"""v2v_b200 — B200-native video-to-voxel hot path of HYLZ-2019/V2V.

Host layer (Python) above the C ABI of ``lib/libv2v_b200.so`` (include/v2v_b200.h).
Importing the package does not need a GPU; every compute call does and fails
loudly without one (no CPU fallback).
"""
from . import _lib
from ._lib import V2VError, launch_count
from .esim import (EventEmulator, EsimOutput, esim_log_lut, frames_to_voxel, draw_reference_randomness,
                   philox_fields)
from .events import (MakeVoxelMixin, event_count_map, events_to_image, events_to_image_torch,
                     events_to_neg_pos_voxel_torch, events_to_voxel, events_to_voxel_torch, fps_window_offsets, make_voxel,
                     pack_events_n5, voxelize_windows)
from .datasets import (ImgsToVoxelsMixin, V2VVoxelizer, bgr_to_gray, degrade_value_map, sample_pause_indices,
                       sample_v2e_params)
from .pipeline import HostPipeline
from .consumer import bin_abs_sums, normalize_batch_voxel, put_accumulate, put_accumulate_bins, voxel_value_hist
from .augment import add_hot_pixels_to_voxels, add_noise_to_voxel, cached_sequence_item

__all__ = [
    "V2VError", "launch_count", "bgr_to_gray", "degrade_value_map", "sample_pause_indices", "EventEmulator", "EsimOutput", "esim_log_lut", "frames_to_voxel",
    "draw_reference_randomness", "philox_fields", "MakeVoxelMixin", "event_count_map", "events_to_image",
    "events_to_image_torch", "events_to_neg_pos_voxel_torch", "events_to_voxel", "events_to_voxel_torch",
    "make_voxel", "voxelize_windows", "fps_window_offsets", "pack_events_n5", "ImgsToVoxelsMixin", "V2VVoxelizer", "sample_v2e_params", "HostPipeline",
    "normalize_batch_voxel", "bin_abs_sums", "voxel_value_hist", "put_accumulate", "put_accumulate_bins", "add_noise_to_voxel", "add_hot_pixels_to_voxels", "cached_sequence_item",
]
