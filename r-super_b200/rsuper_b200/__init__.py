"""rsuper_b200 — B200-native (sm_100a) kernels for the R-Super 3D segmentation train step.

Host-side mirror of the reference's plugin interface for this path:
  * `rsuper_b200.unet.B200UNet`            <-> rsuper_train/model/dim3/unet.py:UNet
  * `rsuper_b200.medformer.B200MedFormer`  <-> rsuper_train/model/dim3/medformer.py:MedFormer
  * `rsuper_b200.plugin.get_model`         <-> rsuper_train/model/utils.py:get_model
  * `rsuper_b200.losses.calculate_loss`    <-> rsuper_train/training/losses_foundation.py:calculate_loss
  * `rsuper_b200.optim.B200AdamW`          <-> get_optimizer (AdamW) + clip_grad_norm_ + update_ema_variables, train_ddp.py:352-357
  * `rsuper_b200.train_step.B200TrainStep` <-> the body of train_epoch (train_ddp.py:310-357) incl. DDP's gradient all-reduce
The compute path is the C-ABI library librsuper_b200.so (include/rsuper_b200.h); there is no
CPU or PyTorch fallback.
"""
__version__ = "0.2.0"
