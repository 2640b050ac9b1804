"""B200-native IM-MoCo instance optimiser (hot path of multimodallearning/MICCAI24_IMMoCo).

Public surface = the reference's own names for this path:
  imcoco_motion_correction, IMMoCo, make_grids, network_config, mot_network_config,
  encoding_config, ClearCache                        (src/models/immoco.py)
  NetworkWithInputEncoding                           (tinycudann, as used at immoco.py:60-65)
  FFT, IFFT                                          (src/utils/data_utils.py:29-34)
  GradientEntropyLoss                                (src/utils/losses.py:20-40)
  extract_movement_groups                            (src/utils/motion_utils.py:56-109)
  make_test_set / load_test_set / run_test_immoco    (src/utils/prepareData.py:144-216, src/test/test_immoco.py:27-130)
Every compute call goes to hand-written sm_100a kernels in libimmoco_b200.so through the C ABI of
include/immoco_b200.h; there is no CPU or eager-PyTorch fallback.
"""
from ._native import build, lib, LIB_PATH, EXPORTED_SYMBOLS  # noqa: F401
from .immoco import (ClearCache, FitEngine, IMMoCo, LineStructure, clear_caches, encoding_config,  # noqa: F401
                     imcoco_motion_correction, lambda_schedule, make_grids, mot_network_config,
                     network_config, run_batched)
from .autofocusing import Autofocusing, autofocus_motion_correction  # noqa: F401
from .batch import reconstruct_batch  # noqa: F401
from .kld_net import Unet, detect_motion_lines, get_unet, kld_net_input, movement_masks_from_kspace  # noqa: F401
from .metrics import calmetric2D, crop_metrics, my_psnr, normalize, rmse  # noqa: F401
from .motion_utils import (extract_movement_groups, generate_list, get_rand_int, lines_from_mask,  # noqa: F401
                           motion_simulation2D, rotation_matrix_2d)
from .ops import FFT, IFFT, GradientEntropyLoss, NetworkWithInputEncoding  # noqa: F401
from .sharding import gather_images, reconstruct_slices, shard_indices  # noqa: F401
from .evaluation import (evaluate_scenarios, load_test_set, make_test_set, run_test_immoco, save_test_set,  # noqa: F401
                         summarize_metrics, validate_test_set)

__all__ = [
    "imcoco_motion_correction", "IMMoCo", "make_grids", "network_config", "mot_network_config",
    "encoding_config", "ClearCache", "NetworkWithInputEncoding", "FFT", "IFFT",
    "GradientEntropyLoss", "extract_movement_groups", "lines_from_mask", "FitEngine",
    "LineStructure", "lambda_schedule", "run_batched", "clear_caches", "build", "lib", "reconstruct_batch", "reconstruct_slices",
    "gather_images", "shard_indices", "calmetric2D", "crop_metrics", "my_psnr", "normalize", "rmse",
    "Autofocusing", "autofocus_motion_correction", "get_unet", "Unet", "kld_net_input", "detect_motion_lines", "movement_masks_from_kspace",
    "motion_simulation2D", "generate_list", "get_rand_int", "rotation_matrix_2d",
    "make_test_set", "save_test_set", "load_test_set", "validate_test_set", "run_test_immoco", "summarize_metrics",
    "evaluate_scenarios",
]
