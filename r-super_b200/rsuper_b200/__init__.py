"""rsuper_b200 — B200-native (sm_100a) kernels for the R-Super 3D segmentation train step.

Host-side mirror of the reference's plugin interface for this path:
  * `rsuper_b200.unet.B200UNet`            <-> rsuper_train/model/dim3/unet.py:UNet
  * `rsuper_b200.plugin.get_model`         <-> rsuper_train/model/utils.py:get_model
  * `rsuper_b200.losses.calculate_loss`    <-> rsuper_train/training/losses_foundation.py:calculate_loss
The compute path is the C-ABI library librsuper_b200.so (include/rsuper_b200.h); there is no
CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
