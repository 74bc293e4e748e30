"""ccst_b200 -- B200-native AdaIN style-transfer hot path of JeremyCJM/CCST.

Operator surface (same names as the reference):

    from ccst_b200 import (calc_mean_std, adaptive_instance_normalization,
                           adaIN_StyleStat_ContentFeat, calc_sum, style_transfer)
    from ccst_b200 import net            # net.vgg / net.decoder nn.Sequential definitions
    from ccst_b200.overall import OverallStyleAccumulator

Everything executes in libccst_b200.so (hand-written sm_100a CUDA behind a C
ABI, `include/ccst_b200.h`); importing the package does not need a GPU, calling
an operator without one raises.
"""
from .function import (EPS, WelfordState, adaIN_StyleStat_ContentFeat, adain_blend,
                       adaptive_instance_normalization, calc_mean_std, calc_mean_std_batch,
                       calc_mean_std_vector, calc_sum, mixstyle, mixstyle_stats, mse_loss)
from .transfer import (Engine, engine_for, resize, resize_input_u8, save_image_quantize, style_transfer,
                       style_transfer_u8, to_tensor_u8)

__all__ = [
    "EPS", "WelfordState", "adaIN_StyleStat_ContentFeat", "adain_blend",
    "adaptive_instance_normalization", "calc_mean_std", "calc_mean_std_batch", "calc_mean_std_vector", "calc_sum",
    "mse_loss", "mixstyle", "mixstyle_stats",
    "Engine", "engine_for", "style_transfer", "style_transfer_u8", "to_tensor_u8", "save_image_quantize", "resize", "resize_input_u8",
]
