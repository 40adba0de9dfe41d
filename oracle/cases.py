"""Named parity cases shared by the golden generator and the tests (TEST INFRASTRUCTURE)."""

SWIN_CASES = {
    # every pad path: grid 13x16x12 -> 18x21x12 | 7x8x6 -> 12x14x6 | 4x4x3 -> 6x7x6 | 2x2x2 -> 6x7x6
    "swin_tiny_odd": dict(embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8], window_size=[6, 7, 6],
                          patch_size=[4, 4, 4], num_classes=5, input=[2, 1, 50, 61, 47], drop_path=0.1),
    # no padding anywhere in stage 0, one window in stage 1
    "swin_small_even": dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=[6, 7, 6],
                            patch_size=[4, 4, 4], num_classes=3, input=[1, 1, 48, 56, 48], drop_path=0.0),
}

VIT_CASES = {
    "vit_tiny": dict(embed_dim=128, depth=2, num_heads=2, patch_size=[16, 16, 16], img_size=[32, 48, 32],
                     mlp_ratio=4.0, num_classes=3, input=[2, 1, 32, 48, 32]),
    # 5x4x4 patches + cls = 81 tokens: more than one 64-wide key tile, ragged tail
    "vit_small": dict(embed_dim=128, depth=2, num_heads=2, patch_size=[16, 16, 16], img_size=[80, 64, 64],
                      mlp_ratio=4.0, num_classes=5, input=[1, 1, 80, 64, 64]),
}

SWIN_FULL = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=[6, 7, 6],
                 patch_size=[4, 4, 4])
VIT_FULL = dict(embed_dim=384, depth=12, num_heads=6, patch_size=[16, 16, 16], img_size=[144, 160, 144], mlp_ratio=4.0)


def swin_ctor_kwargs(case, drop_path=None):
    import torch
    return dict(in_channels=1, patch_size=case["patch_size"], embed_dim=case["embed_dim"], depths=case["depths"],
                num_heads=case["num_heads"], window_size=case["window_size"], mlp_ratio=4.0, qkv_bias=True,
                dropout=0.0, attention_dropout=0.0,
                stochastic_depth_prob=case.get("drop_path", 0.0) if drop_path is None else drop_path,
                num_classes=case["num_classes"], norm_layer=torch.nn.LayerNorm)


def vit_ctor_kwargs(case):
    return dict(img_size=tuple(case["img_size"]), num_classes=case["num_classes"], in_channels=1,
                patch_size=tuple(case["patch_size"]), mlp_ratio=case["mlp_ratio"], dropout=0.0,
                attention_dropout=0.0, embed_dim=case["embed_dim"], num_heads=case["num_heads"], depth=case["depth"])
