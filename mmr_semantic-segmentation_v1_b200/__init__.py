"""B200-native (sm_100a) segmentation train/infer step of MMR_semantic-segmentation_v1.

Import as `mmrseg_b200`.  Sub-modules:
  _lib       ctypes binding of libmmrseg.so (C-ABI in include/mmrseg.h)
  convplan   K-step tables for the tcgen05 implicit-GEMM convolution
  engine     static execution plan of a whole network (forward, backward, Adam)
  models     UnetPlusPlus / ResNetUNet with the reference's constructors and state_dict keys
  losses     DiceLoss / DiceCELoss / mixed loss (reference signatures)
  metrics    Evaluate / get_stats / iou_score (reference signatures)
"""
__version__ = "0.1.0"
