"""Shadow packages that put the vsn_b200 implementations behind the reference's own import paths.

Put this directory AHEAD of the reference checkout on `PYTHONPATH` (the reference's trainer appends its own
root to `sys.path` last, train/train_transformer.py:51):

    PYTHONPATH=<repo>:<repo>/vit-stability-neurodegeneration_b200/dropin:<reference> python train/train_transformer.py ...

`models`, `regularization` and `utils` then resolve here; each extends its `__path__` with the reference's
package directory, so every module that is not replaced (label smoothing, helpers, samplers, MedViT, ResNet...)
still comes from the reference, unmodified.  Replaced: models.swin_transformer_3d, models.vit_3d,
regularization.sam, utils.ema.
"""
