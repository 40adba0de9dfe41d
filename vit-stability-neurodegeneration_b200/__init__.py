"""vsn_b200 — B200-native (sm_100a) training hot path for the 3D Swin / ViT classifiers of
EloiNavet/ViT-Stability-Neurodegeneration.

Layout
  csrc/        CUDA kernels + the C ABI (`include/vsn_b200.h`), built in-tree into libvsn_b200.so
  _lib.py      ctypes binding of that ABI (fails loudly if the library or a B200 is missing)
  ops.py       thin tensor-level wrappers (allocation + argument marshalling only)
  swin.py      Swin-3D block / stage / model functions with hand-written backward
  vit.py       ViT-3D equivalents
  optim.py     SAM, EMA, FusedAdamW on the multi-tensor kernels
  ddp.py       bucketed gradient all-reduce for one-process-per-GPU data parallelism
  train.py     one optimiser step (in-place gradient arena, CUDA-graph replay, fused micro-batches, SAM / EMA)
  data.py      the input step on the device: MixUp plan + mix + whole-image z-score of the fp16 volumes
  tta.py       test-time augmentation (all views in one launch) + snapshot ensemble
  dropin/      `models/`, `regularization/`, `utils/` packages that shadow the reference's modules so
               train/train_transformer.py and eval/eval_transformer.py run unchanged on this path
"""
__version__ = "0.1.0"
