"""Backbone / segmentor kwargs of the shipped CDSegNet configs, restated as plain dicts so that
bench.py and the tests can build the exact architectures without the reference's Config loader.
Source: configs/scannet/CDSegNet.py:55-141 (ScanNet), configs/nuscenes/CDSegNet.py:29-30, 55-141 (nuScenes:
same backbone with 4 input channels and 16 classes), configs/scannet200/CDSegNet.py (200 classes)."""

ORDER4 = ("z", "z-trans", "hilbert", "hilbert-trans")


def backbone_cfg(in_channels=6, num_classes=20, condition=True, patch=1024, shuffle_orders=True, enable_flash=True):
    return dict(
        c_in_channels=in_channels, n_in_channels=in_channels, order=ORDER4,
        c_stride=(4, 4), c_enc_depths=(2, 2, 2), c_enc_channels=(32, 64, 128), c_enc_num_head=(2, 4, 8),
        c_enc_patch_size=(patch,) * 3, c_dec_depths=(2, 2), c_dec_channels=(64, 64), c_dec_num_head=(4, 4),
        c_dec_patch_size=(patch,) * 2,
        n_stride=(2, 2, 2, 2), n_enc_depths=(2, 2, 2, 6, 6), n_enc_channels=(32, 64, 128, 256, 512),
        n_enc_num_head=(2, 4, 8, 16, 32), n_enc_patch_size=(patch,) * 5, n_dec_depths=(2, 2, 2, 2),
        n_dec_channels=(64, 64, 128, 256), n_dec_num_head=(4, 4, 8, 16), n_dec_patch_size=(patch,) * 4,
        mlp_ratio=4, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0, drop_path=0.3,
        shuffle_orders=shuffle_orders, pre_norm=True, enable_rpe=False, enable_flash=enable_flash,
        upcast_attention=False, upcast_softmax=False, cls_mode=False, pdnorm_bn=False, pdnorm_ln=False,
        pdnorm_decouple=True, pdnorm_adaptive=False, pdnorm_affine=True,
        pdnorm_conditions=("ScanNet", "S3DIS", "Structured3D"),
        num_classes=num_classes, T_dim=128, tm_bidirectional=False, tm_feat=1.0, tm_restomer=False,
        condition=condition, skip_connection_mode="cat", b_factor=[1.0] * 4, s_factor=[1.0] * 4,
        skip_connection_scale=True, skip_connection_scale_i=False,
    )


def segmentor_cfg(in_channels=6, num_classes=20, condition=True, **kw):
    """configs/scannet/CDSegNet.py:55-141 (model=dict(type="DefaultSegmentorV2", ...))."""
    return dict(
        type="DefaultSegmentorV2",
        backbone=dict(type="PT-v3m1", **backbone_cfg(in_channels, num_classes, condition, **kw)),
        criteria=None, loss_type="GLS", task_num=2, num_classes=num_classes, T=1000, beta_start=0, beta_end=1000,
        noise_schedule="cosine", T_dim=128, dm=True, dm_input="xt", dm_target="noise", dm_min_snr=None,
        condition=condition, c_in_channels=in_channels,
    )
