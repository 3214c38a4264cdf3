"""Drop-in heads of the hot path, registered under the reference's names in ``HEADS``.

* UncertPropPnPOptimizer  <- monorun/models/roi_heads/bbox_3d_heads/optimizers/uncert_prop_pnp_optimizer.py:12-99
* FCNNOCDecoder           <- monorun/models/roi_heads/bbox_3d_heads/dense_decoders/fcn_noc_decoder.py:15-267
* UncertProjectionHead    <- .../reprojection_heads/uncert_projection_head.py (test-time parts: get_distance, coder)
* MonoRUnRoIHead          <- monorun/models/roi_heads/monorun_roi_head.py:13-40, simple_test :442-605 (hot sequence
                             :509-534), result packaging :607-655
* SingleRoIExtractor      <- mmdet.models.roi_heads.roi_extractors.SingleRoIExtractor as configured at
                             configs/kitti_multiclass.py:38-43, 83-88 (RoIAlign per FPN level; a caller-side stage)
* FCExtractor[MonteCarlo] <- .../global_extractors/fc_extractor.py:12-156, fc_extractor_monte_carlo.py:21-82 (the caller-side
                             stage that produces latent / dimensions / reg_fc_out; torch Linear layers, not a kernel of
                             this path)

Constructor kwargs, attribute names, state-dict keys and return tuples follow the reference so that the
``roi_head`` blocks of configs/kitti_*.py build and pretrained weights would load.  Training-only members
(losses, target builders) are accepted and ignored: the PnP forward is non-differentiable in the reference too
(pnp_uncert.py:33) and training is out of this path's scope.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .coders import coords_2d_from_rois
from .registry import (HEADS, ROI_EXTRACTORS, build_coord_coder, build_dim_coder, build_head, build_pnp,
                       build_proj_error_coder, build_roi_extractor, build_rotation_coder)


@HEADS.register_module()
class UncertPropPnPOptimizer(nn.Module):
    """Pose head: log-std -> weights, clip ranges, PnP, covariance calibration.

    The reference permutes the three NCHW maps to (N, 784, C) and calls the op (:82-95); here the NCHW tensors and
    the log-std go straight into the kernel (``PnPUncert.forward_dense``), which fuses :73 and :82-84.
    """

    def __init__(self, loss_rot=None, loss_trans=None, loss_calib=None,
                 rotation_coder=dict(type='Vec2DRotationCoder'),
                 pnp=dict(type='PnPUncert', z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True,
                          forward_exact_hessian=False),
                 allowed_border=200, epnp_ransac_thres_ratio=0.2, std_scale=10):
        super(UncertPropPnPOptimizer, self).__init__()
        pnp = dict(pnp)
        pnp.pop('backward_exact_hessian', None)  # present in the reference's default dict, rejected by PnPUncert
        self.pnp = build_pnp(pnp)
        self.epnp_ransac_thres_ratio = epnp_ransac_thres_ratio
        self.allowed_border = allowed_border
        self.std_scale = std_scale
        self.rotation_coder = build_rotation_coder(rotation_coder)
        self.fp16_enabled = False
        self.loss_rot = self.loss_trans = self.loss_calib = None  # training only
        self.cov_calib_logscale = nn.Parameter(torch.full((4, ), 0, dtype=torch.float))

    def init_weights(self):
        pass

    def forward(self, coords_2d, coords_2d_logstd, coords_3d, cam_intrinsic, img_shapes, init_pose=None):
        """uncert_prop_pnp_optimizer.py:50-99.

        Args:
            coords_2d (Tensor): (Nbatch, 2, h, w);  coords_2d_logstd (Tensor): (Nbatch, 2, h, w)
            coords_3d (Tensor): (Nbatch, 3, h, w);  cam_intrinsic (Tensor): (Nbatch, 3, 3) or (1, 3, 3)
            img_shapes (Tensor): (Nbatch, 2) or (1, 2), rows are (h, w)
        Returns:
            ret_val (Nbatch,) bool, yaw_pred (Nbatch, 1), t_vec_pred (Nbatch, 3),
            pose_cov_pred (Nbatch, 4, 4), pose_cov_calib (Nbatch, 4, 4)
        """
        epnp_ransac_thres = None
        if self.epnp_ransac_thres_ratio is not None and coords_2d.size(0):   # :86-88
            roi_heights = coords_2d[:, 1, -1, 0] - coords_2d[:, 1, 0, 0]
            epnp_ransac_thres = self.epnp_ransac_thres_ratio * roi_heights
        ret_val, yaw_pred, t_vec_pred, pose_cov_pred, _ = self.pnp.forward_dense(
            coords_2d, coords_2d_logstd, coords_3d, cam_intrinsic, self._uv_range(img_shapes), self.std_scale,
            init_pose=init_pose, epnp_ransac_thres=epnp_ransac_thres)
        return ret_val, yaw_pred, t_vec_pred, pose_cov_pred, self._calibrate(pose_cov_pred)

    def _uv_range(self, img_shapes):
        b = float(self.allowed_border)
        uv_range = img_shapes.new_empty((img_shapes.size(0), 4), dtype=torch.float32)
        uv_range[:, 0] = -b                      # :75-80
        uv_range[:, 1] = img_shapes[:, 1] + b
        uv_range[:, 2] = -b
        uv_range[:, 3] = img_shapes[:, 0] + b
        return uv_range

    def _calibrate(self, pose_cov_pred):
        cov_calib_scale = torch.exp(self.cov_calib_logscale)
        return (cov_calib_scale * cov_calib_scale[:, None]) * pose_cov_pred  # :96-97

    def forward_fused(self, noc_pred, proj_logstd, rois, dimensions, dimensions_var, cam_intrinsic, img_shapes,
                      coord_coder, proj_error_coder, distance=None, init_pose=None, labels=None, num_classes=0,
                      dim_coder=None, dim_labels=None):
        """Same five outputs as :meth:`forward`, from the dense head's RAW maps: NOC decode, variance-propagated
        log-std and the RoI pixel grid run inside the PnP kernel (one launch for monorun_roi_head.py:513-529).
        With ``dim_coder`` the dimensions come in encoded, are decoded by the same launch (:503-507) and the decoded
        ``dimensions``, ``dimensions_var`` are appended to the outputs."""
        ret_val, yaw_pred, t_vec_pred, pose_cov_pred, _, *decoded = self.pnp.forward_fused(
            noc_pred, proj_logstd, rois, dimensions, dimensions_var, cam_intrinsic, self._uv_range(img_shapes),
            self.std_scale, coord_coder, proj_error_coder, distance=distance, init_pose=init_pose, labels=labels,
            num_classes=num_classes, ransac_ratio=self.epnp_ransac_thres_ratio or 0.0, dim_coder=dim_coder,
            dim_labels=dim_labels)
        return (ret_val, yaw_pred, t_vec_pred, pose_cov_pred, self._calibrate(pose_cov_pred), *decoded)


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule with norm_cfg=None: Conv2d(bias=True) + ReLU (fcn_noc_decoder.py:98-105)."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0):
        super(ConvModule, self).__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, padding=padding)
        self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.activate(self.conv(x))


class CARAFEPack(nn.Module):
    """Functional restatement of mmcv.ops.carafe.CARAFEPack (source not in this tree):
    channel_compressor 1x1 -> content_encoder k_enc x k_enc -> pixel_shuffle -> softmax over the k_up^2 taps ->
    per-pixel reassembly of the k_up x k_up neighbourhood of the low-resolution map."""

    def __init__(self, channels, scale_factor, up_kernel=5, up_group=1, encoder_kernel=3, encoder_dilation=1,
                 compressed_channels=64):
        super(CARAFEPack, self).__init__()
        self.channels, self.scale_factor, self.up_kernel, self.up_group = channels, scale_factor, up_kernel, up_group
        self.channel_compressor = nn.Conv2d(channels, compressed_channels, 1)
        self.content_encoder = nn.Conv2d(
            compressed_channels, up_kernel * up_kernel * up_group * scale_factor * scale_factor, encoder_kernel,
            padding=int((encoder_kernel - 1) * encoder_dilation / 2), dilation=encoder_dilation)

    def init_weights(self):
        nn.init.xavier_uniform_(self.channel_compressor.weight)
        nn.init.constant_(self.channel_compressor.bias, 0)
        nn.init.normal_(self.content_encoder.weight, std=0.001)
        nn.init.constant_(self.content_encoder.bias, 0)

    def forward(self, x):
        n, c, h, w = x.shape
        s, k, g = self.scale_factor, self.up_kernel, self.up_group
        mask = self.content_encoder(self.channel_compressor(x))
        mask = F.pixel_shuffle(mask, s)                                   # (n, g*k*k, h*s, w*s)
        mask = F.softmax(mask.view(n, g, k * k, h * s, w * s), dim=2)
        patches = F.unfold(x, k, padding=k // 2).view(n, g, c // g, k * k, h, w)
        patches = patches.repeat_interleave(s, dim=4).repeat_interleave(s, dim=5)  # nearest source pixel
        out = (patches * mask.unsqueeze(2)).sum(dim=3)
        return out.view(n, c, h * s, w * s)


@HEADS.register_module()
class FCNNOCDecoder(nn.Module):
    """Dense correspondence head, fcn_noc_decoder.py:15-267 (forward :189-240, slice_pred :242-267)."""

    def __init__(self, num_convs=3, roi_feat_size=14, in_channels=256, conv_kernel_size=3, conv_out_channels=256,
                 num_classes=3, class_agnostic=False,
                 upsample_cfg=dict(type='carafe', scale_factor=2, up_kernel=5, up_group=1, encoder_kernel=3,
                                   encoder_dilation=1, compressed_channels=64),
                 num_convs_upsampled=1, conv_cfg=None, norm_cfg=None, loss_noc=None, noc_channels=3,
                 uncert_channels=2, dropout2d_rate=0.2, num_dropout2d_layers=1, flip_correction=True, plugins=None,
                 coord_coder=dict(type='NOCCoder', target_means=(-0.1, -0.5, 0.0), target_stds=(0.35, 0.23, 0.34),
                                  eps=1e-5),
                 use_latent_vec=True, latent_activation=None, latent_channels=16):
        super(FCNNOCDecoder, self).__init__()
        assert num_convs > 0 and conv_cfg is None and norm_cfg is None and plugins is None
        up = dict(type='carafe', scale_factor=2, up_kernel=5, up_group=1, encoder_kernel=3, encoder_dilation=1,
                  compressed_channels=64)  # mmcv CARAFEPack defaults, partially overridden by the config (:98)
        up.update(upsample_cfg)
        if up['type'] != 'carafe':
            raise NotImplementedError('only the carafe upsampler used by every reference config is provided')
        self.num_convs, self.num_convs_upsampled = num_convs, num_convs_upsampled
        self.in_channels, self.conv_out_channels = in_channels, conv_out_channels
        self.num_classes, self.class_agnostic = num_classes, class_agnostic
        self.upsample_method, self.scale_factor = up.pop('type'), up.pop('scale_factor')
        self.fp16_enabled = False
        self.loss_noc = None  # training only
        self.flip_correction = flip_correction
        self.noc_channels, self.uncert_channels = noc_channels, uncert_channels
        self.channel_per_class = noc_channels + uncert_channels
        self.coord_coder = build_coord_coder(coord_coder)
        self.use_latent_vec = use_latent_vec
        self.latent_activation = (nn.ReLU() if latent_activation == 'ReLU'
                                  else nn.LeakyReLU() if latent_activation == 'LeakyReLU' else None)
        if use_latent_vec:
            self.latent_decoder = nn.Linear(latent_channels, conv_out_channels)
        pad = (conv_kernel_size - 1) // 2
        self.convs = nn.ModuleList(
            [ConvModule(in_channels if i == 0 else conv_out_channels, conv_out_channels, conv_kernel_size, pad)
             for i in range(num_convs)])
        self.upsample = CARAFEPack(channels=conv_out_channels, scale_factor=self.scale_factor, **up)
        self.convs_upsampled = nn.ModuleList(
            [ConvModule(conv_out_channels, conv_out_channels, conv_kernel_size, pad)
             for _ in range(num_convs_upsampled)])
        final_out = self.channel_per_class * (1 if class_agnostic else num_classes) * (2 if flip_correction else 1)
        self.conv_final = nn.Conv2d(conv_out_channels, final_out, 1)
        self.use_dropout2d = dropout2d_rate > 0
        if self.use_dropout2d:
            self.dropout2d = nn.Dropout2d(dropout2d_rate)
        self.num_dropout2d_layers = num_dropout2d_layers

    def init_weights(self):
        self.upsample.init_weights()
        nn.init.kaiming_normal_(self.conv_final.weight, mode='fan_out', nonlinearity='relu')
        nn.init.constant_(self.conv_final.bias, 0)
        if self.use_latent_vec:
            nn.init.constant_(self.latent_decoder.weight, 0)
            nn.init.constant_(self.latent_decoder.bias, 0)

    def forward(self, x, latent_pred, latent_var, labels, flip=False):
        noc_pred, noc_var, proj_logstd = self.slice_pred(self.forward_all(x, latent_pred, flip), labels)
        return noc_pred, noc_var, proj_logstd, None

    def forward_all(self, x, latent_pred, flip=False, native=False):
        """fcn_noc_decoder.py:189-235 up to (not including) slice_pred: the unsliced ``all_pred`` [N, 5*C, 2h, 2w]
        (a strided view of the flip-paired conv output).  The fused PnP entry slices it by class itself.

        ``native=True`` runs the layers in libmonorun_head.so (tcgen05 implicit-GEMM convolutions + CARAFE kernel,
        bf16 operands / fp32 accumulation, inference only) instead of the fp32 torch modules below."""
        if native and x.size(0) > 0:
            from .dense_head import DenseHeadB200
            runner = getattr(self, '_b200_runner', None)
            if runner is None or runner.device != x.device:
                runner = DenseHeadB200(self)      # packs the weights once; call drop_native_cache() after changing them
                object.__setattr__(self, '_b200_runner', runner)
            all_pred = runner.forward(x, latent_pred)
            return self._select_flip_half(all_pred, flip)
        if self.use_dropout2d and self.num_dropout2d_layers > 0:
            x = self.dropout2d(x)
        for i, conv in enumerate(self.convs):
            x = conv(x)
            if self.use_dropout2d and i + 1 < self.num_dropout2d_layers:
                x = self.dropout2d(x)
        if self.use_latent_vec:
            if self.latent_activation is not None:
                latent_pred = self.latent_activation(latent_pred)
            x = x + self.latent_decoder(latent_pred)[..., None, None]
        if x.size(0) == 0:
            c = self.conv_final.out_channels // (2 if self.flip_correction else 1)
            s = x.size(2) * self.scale_factor
            all_pred = x.new_zeros((0, c, s, s))
        else:
            x = self.upsample(x)
            for conv_upsampled in self.convs_upsampled:
                x = conv_upsampled(x)
            all_pred = self._select_flip_half(self.conv_final(x), flip)
        return all_pred

    def drop_native_cache(self):
        object.__setattr__(self, '_b200_runner', None)

    def _select_flip_half(self, all_pred, flip):
        if self.flip_correction:   # :225-235
            all_pred = all_pred.view(all_pred.size(0), 2, all_pred.size(1) // 2, all_pred.size(2), all_pred.size(3))
            if isinstance(flip, bool):
                all_pred = all_pred[:, 1 if flip else 0]
            else:
                inds = torch.arange(0, all_pred.size(0), dtype=torch.long, device=all_pred.device)
                all_pred = all_pred[inds, inds.new_tensor(flip)]
        return all_pred

    def slice_pred(self, all_pred, labels):
        k = 1 if self.class_agnostic else self.num_classes
        all_noc_pred, all_proj_logstd = all_pred.split([self.noc_channels * k, self.uncert_channels * k], dim=1)
        if self.class_agnostic:
            return all_noc_pred, None, all_proj_logstd
        n, _, h, w = all_noc_pred.size()
        inds = torch.arange(0, n, dtype=torch.long, device=all_noc_pred.device)
        noc_pred = all_noc_pred.view(n, self.num_classes, 3, h, w)[inds, labels]
        proj_logstd = all_proj_logstd.view(n, self.num_classes, self.uncert_channels, h, w)[inds, labels]
        return noc_pred, None, proj_logstd


@HEADS.register_module()
class UncertProjectionHead(nn.Module):
    """Test-time members of uncert_projection_head.py: the projection-error coder and get_distance (:104-109)."""

    def __init__(self, loss_proj=None, z_min=0.5, allowed_border=200,
                 proj_error_coder=dict(type='DistanceInvarProjErrorCoder', ref_length=1.6, ref_focal_y=722,
                                       target_std=0.15),
                 distance_mode='range'):
        super(UncertProjectionHead, self).__init__()
        assert distance_mode in ['z-depth', 'range']
        self.z_min, self.allowed_border, self.distance_mode = z_min, allowed_border, distance_mode
        self.proj_error_coder = build_proj_error_coder(proj_error_coder)

    def get_distance(self, t_vec):
        return t_vec[:, 2] if self.distance_mode == 'z-depth' else torch.norm(t_vec, p=2, dim=1)


class BatchNormSmooth1D(nn.modules.batchnorm._NormBase):
    """mlp_score_head.py:142-184: batch norm that ALWAYS normalises with the running statistics (updated first when
    training).  Inference only here: ``forward`` is the eval branch (:177-178)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super(BatchNormSmooth1D, self).__init__(num_features, eps, momentum, affine, track_running_stats)

    def _check_input_dim(self, input):
        if input.dim() != 2 and input.dim() != 3:
            raise ValueError('expected 2D or 3D input (got {}D input)'.format(input.dim()))

    def forward(self, input):
        self._check_input_dim(input)
        if self.training:
            raise RuntimeError('BatchNormSmooth1D: inference only in this repo (call .eval())')
        return input.sub(self.running_mean).div((self.running_var + self.eps).sqrt()).mul(self.weight).add(self.bias)


@HEADS.register_module()
class MLPScoreHead(nn.Module):
    """Score head (mlp_score_head.py:11-115), inference side: same constructor kwargs, sub-module names and
    state-dict keys (``pose_norm``, ``pose_fcs``, ``fused_fcs``, ``fc_out``).

    ``forward`` is the reference's torch sequence (and the fp32 reference of the tests).  ``forward_rows`` is the
    B200 path: ONE launch (``mrpnp_score_stage``) builds the normalised 17-feature rows straight from the solver's result
    rows (covariance calibration, test-time correction, lower triangle, concatenation, pose_norm), runs the three Linear
    layers and finishes (sigmoid, invalid -> 0, product with the 2-D score, the [l,h,w,x,y,z,ry,score] rows).  Other
    network shapes (more layers, fusion 'concat') fall back to ``mrpnp_pose_features`` + library GEMMs +
    ``mrpnp_finish_scores``."""

    def __init__(self, reg_fc_out_channels=1024, num_pose_fcs=1, pose_fc_out_channels=1024, fusion_type='add',
                 num_fused_fcs=1, fc_out_channels=256, loss_score=None, mode='linear_average', iou_thres=0.7,
                 linear_coefs=(-0.5, 2), detach_preds=True, use_pose_norm=True, train_cfg=None):
        super(MLPScoreHead, self).__init__()
        assert mode in ['average', 'thres', 'linear_average'] and fusion_type in ['add', 'concat']
        assert num_pose_fcs > 0 and num_fused_fcs > 0
        self.num_pose_fcs, self.num_fused_fcs = num_pose_fcs, num_fused_fcs
        self.fc_out_channels, self.pose_fc_out_channels = fc_out_channels, pose_fc_out_channels
        self.reg_fc_out_channels, self.fusion_type, self.use_pose_norm = reg_fc_out_channels, fusion_type, use_pose_norm
        self.mode, self.iou_thres, self.linear_coefs, self.detach_preds = mode, iou_thres, linear_coefs, detach_preds
        self.loss_score, self.train_cfg = loss_score, train_cfg   # training is out of scope: kept as configuration
        self.fp16_enabled = False
        self.pre_sigmoid = True
        self.relu = nn.ReLU(inplace=True)
        layer_dim = reg_fc_out_channels
        if fusion_type == 'add':
            assert pose_fc_out_channels == reg_fc_out_channels
        else:
            layer_dim += pose_fc_out_channels
        pose_last_layer_dim = 1 + 3 + 10 + 3  # only the lower triangle of the covariance
        if use_pose_norm:
            self.pose_norm = BatchNormSmooth1D(pose_last_layer_dim, momentum=0.01)
        self.pose_fcs = nn.ModuleList(nn.Linear(pose_last_layer_dim if i == 0 else pose_fc_out_channels,
                                                pose_fc_out_channels) for i in range(num_pose_fcs))
        self.fused_fcs = nn.ModuleList(nn.Linear(layer_dim if i == 0 else fc_out_channels, fc_out_channels)
                                       for i in range(num_fused_fcs))
        self.fc_out = nn.Linear(fc_out_channels, 1)

    def init_weights(self):
        for m in list(self.pose_fcs) + list(self.fused_fcs):
            nn.init.xavier_uniform_(m.weight)
            nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.fc_out.weight, 0, 0.01)
        nn.init.constant_(self.fc_out.bias, 0)

    def _mlp(self, x, reg_fc_out):
        for fc in self.pose_fcs:
            x = self.relu(fc(x))
        x = x + reg_fc_out if self.fusion_type == 'add' else torch.cat([x, reg_fc_out], dim=1)
        for fc in self.fused_fcs:
            x = self.relu(fc(x))
        return self.fc_out(x).squeeze(1)

    def forward(self, reg_fc_out, yaw, t_vec, pose_cov, dimensions):
        """mlp_score_head.py:94-115 (raw scores, shape (n,))."""
        cov_x_inds, cov_y_inds = torch.tril_indices(4, 4, device=pose_cov.device)
        x = torch.cat([yaw, t_vec, pose_cov[:, cov_x_inds, cov_y_inds], dimensions], dim=1)
        if self.use_pose_norm:
            x = self.pose_norm(x)
        return self._mlp(x, reg_fc_out)

    def forward_rows(self, reg_fc_out, rows, dimensions, cov_calib_logscale=None, cov_correction_sd=0.0,
                     distance_z_depth=False, calib_scoring=False, det_scores=None, native_mlp=True):
        """Solver result rows [N,24] -> (scores [N], bbox_3d [N,8], pose_cov_calib [N,4,4]); the test-time tail of
        MonoRUnRoIHead.simple_test (monorun_roi_head.py:530-556): one launch (mrpnp_score_stage) for image-sized batches of
        the reference configs' network shape, else two launches around the library GEMMs."""
        from . import pnp
        norm = self.pose_norm if self.use_pose_norm else None
        # (streams the fused layer's weights once per 8 objects: the right shape for an image's <= 100 RoIs; from a few
        # hundred objects on the library GEMMs below are faster -- 1024 objects: 0.16 ms against 0.09 ms)
        if native_mlp and rows.shape[0] <= self.native_mlp_max_objects and self.num_pose_fcs == 1 \
                and self.num_fused_fcs == 1 and self.fusion_type == 'add':
            # every reference config: the whole stage is ONE launch (mrpnp_score_stage)
            w = self._native_weights()
            scores, bbox_3d, cov_calib, _ = pnp.score_stage(
                rows, dimensions, reg_fc_out, *w, cov_calib_logscale=cov_calib_logscale, cov_correction_sd=cov_correction_sd,
                distance_z_depth=distance_z_depth, use_calib=calib_scoring, pose_norm=norm, det_scores=det_scores,
                pre_sigmoid=self.pre_sigmoid)
            return scores, bbox_3d, cov_calib.view(-1, 4, 4)
        feat, cov_calib = pnp.pose_features(rows, dimensions, cov_calib_logscale, cov_correction_sd, distance_z_depth,
                                            calib_scoring, norm)
        logits = self._mlp(feat, reg_fc_out)
        scores, bbox_3d = pnp.finish_scores(logits, rows, dimensions, det_scores, self.pre_sigmoid)
        return scores, bbox_3d, cov_calib.view(-1, 4, 4)

    native_mlp_max_objects = 384

    def _native_weights(self):
        """fp32 contiguous weights in the layout of ``mrpnp_score_stage`` (fused layer transposed); rebuilt when a
        parameter changes (``_version`` counts in-place updates such as load_state_dict / optimizer steps)."""
        ps = [self.pose_fcs[0].weight, self.pose_fcs[0].bias, self.fused_fcs[0].weight, self.fused_fcs[0].bias,
              self.fc_out.weight, self.fc_out.bias]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        cache = getattr(self, '_b200_weights', None)
        if cache is None or cache[0] != key:
            with torch.no_grad():
                w = (ps[0].detach().float().contiguous(), ps[1].detach().float().contiguous(),
                     ps[2].detach().float().t().contiguous(), ps[3].detach().float().contiguous(),
                     ps[4].detach().float().reshape(-1).contiguous(), ps[5].detach().float().reshape(1).contiguous())
            cache = (key, w)
            object.__setattr__(self, '_b200_weights', cache)
        return cache[1]


@HEADS.register_module()
class FCExtractor(nn.Module):
    """Global extractor (fc_extractor.py:12-156): RoI feature (N, 256, 7, 7) -> [dimensions | latent vector] per class
    and the 1024-d feature the score head reuses.  Same kwargs and state-dict keys (``fcs.i``, ``fc_reg``); the loss
    config is accepted and ignored (training is out of scope)."""

    def __init__(self, with_dim=True, with_latent_vec=True, latent_channels=16, num_fcs=2, in_channels=256,
                 fc_out_channels=1024, num_classes=3, roi_feat_size=7, latent_class_agnostic=False, loss_dim=None,
                 dim_coder=dict(type='MultiClassNormDimCoder'), dropout_rate=0.5, dropout2d_rate=0.2,
                 num_dropout_layers=2):
        super(FCExtractor, self).__init__()
        assert num_fcs > 0
        self.with_dim, self.with_latent_vec = with_dim, with_latent_vec
        self.dim_dim = 3
        self.latent_channels = latent_channels if with_latent_vec else 0
        rs = (roi_feat_size, roi_feat_size) if isinstance(roi_feat_size, int) else tuple(roi_feat_size)
        self.roi_feat_size, self.roi_feat_area = rs, rs[0] * rs[1]
        self.in_channels, self.fc_out_channels, self.num_classes = in_channels, fc_out_channels, num_classes
        self.latent_class_agnostic = latent_class_agnostic
        self.dim_coder = build_dim_coder(dim_coder)
        self.use_dropout, self.use_dropout2d = dropout_rate > 0, dropout2d_rate > 0
        self.dropout_rate, self.dropout2d_rate = dropout_rate, dropout2d_rate
        self.num_dropout_layers = num_dropout_layers
        self.fcs = nn.ModuleList(
            nn.Linear(in_channels * self.roi_feat_area if i == 0 else fc_out_channels, fc_out_channels)
            for i in range(num_fcs))
        out_dim_reg = self.dim_dim + self.latent_channels
        if not latent_class_agnostic:
            out_dim_reg *= num_classes
        self.fc_reg = nn.Linear(fc_out_channels, out_dim_reg)
        self.mc_dropout = False  # FCExtractorMonteCarlo keeps dropout active at test time

    def init_weights(self):  # fc_extractor.py:85-91
        for m in self.fcs:
            nn.init.xavier_uniform_(m.weight, gain=0.33)
            nn.init.normal_(m.bias, mean=0.02, std=0.04)
        nn.init.normal_(self.fc_reg.weight, 0, 0.001)
        nn.init.constant_(self.fc_reg.bias, 0)

    def _fc_forward(self, x):  # fc_extractor.py:94-105
        active = self.training or self.mc_dropout
        if self.use_dropout2d:
            x = F.dropout2d(x, self.dropout2d_rate, active)
        x = x.flatten(1)
        for i, fc in enumerate(self.fcs):
            x = F.relu(fc(x))
            if self.use_dropout and i < self.num_dropout_layers:
                x = F.dropout(x, self.dropout_rate, active)
        return self.fc_reg(x), x

    def forward(self, x):
        dim_latent_pred, feat = self._fc_forward(x)
        return dim_latent_pred, None, None, None, feat

    def _per_class(self, t, labels):
        if self.latent_class_agnostic:
            return t
        inds = torch.arange(len(labels), device=labels.device)
        return t.view(t.size(0), -1, self.dim_dim + self.latent_channels)[inds, labels]

    def slice_pred(self, dim_latent_pred, dim_latent_var, labels):  # fc_extractor.py:135-146
        dim_pred, latent_pred = self._per_class(dim_latent_pred, labels).split(
            [self.dim_dim, self.latent_channels], dim=1)
        return dim_pred, None, latent_pred, None


@HEADS.register_module()
class FCExtractorMonteCarlo(FCExtractor):
    """MC-dropout global extractor (fc_extractor_monte_carlo.py:21-82): at test time the batch is repeated
    ``num_samples`` times with dropout active; the sample mean / variance give the prediction and its epistemic
    variance (``dimensions_var`` of NOCCoder.decode), the feature is the sample mean."""

    def __init__(self, num_samples=50, dropout_rate=0.5, dropout2d_rate=0.2, **kwargs):
        super(FCExtractorMonteCarlo, self).__init__(dropout_rate=dropout_rate, dropout2d_rate=dropout2d_rate, **kwargs)
        assert self.use_dropout and self.use_dropout2d
        self.num_samples = num_samples
        self.mc_dropout = True

    def forward(self, x):
        if self.training:
            return super(FCExtractorMonteCarlo, self).forward(x)
        n = x.size(0)
        pred, feat = self._fc_forward(x.repeat(self.num_samples, 1, 1, 1))      # :44-47
        var, mean = torch.var_mean(pred.view(self.num_samples, n, -1), dim=0)   # :52-56
        feat = feat.view(self.num_samples, n, -1).mean(dim=0)                   # :60-61
        return mean, var, None, None, feat

    def slice_pred(self, dim_latent_pred, dim_latent_var, labels):  # :65-82
        if self.training:
            return super(FCExtractorMonteCarlo, self).slice_pred(dim_latent_pred, dim_latent_var, labels)
        dim_pred, latent_pred = self._per_class(dim_latent_pred, labels).split(
            [self.dim_dim, self.latent_channels], dim=1)
        dim_var, latent_var = self._per_class(dim_latent_var, labels).split(
            [self.dim_dim, self.latent_channels], dim=1)
        return dim_pred, dim_var, latent_pred, latent_var


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(nn.Module):
    """mmdet's SingleRoIExtractor for the two extractors of the config (configs/kitti_multiclass.py:38-43, 83-88):
    every RoI is pooled from ONE pyramid level, ``floor(log2(sqrt(w h) / finest_scale + 1e-6))`` clamped to the levels,
    with RoIAlign(output_size, sampling_ratio, aligned=True) at that level's stride.  torchvision's ``roi_align`` is the
    operator (mmcv is not a dependency); this stage feeds the path, it is not part of it."""

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super(SingleRoIExtractor, self).__init__()
        cfg = dict(roi_layer)
        assert cfg.pop('type', 'RoIAlign') == 'RoIAlign'
        size = cfg.pop('output_size')
        self.output_size = (size, size) if isinstance(size, int) else tuple(size)
        self.sampling_ratio = cfg.pop('sampling_ratio', 0)
        self.aligned = cfg.pop('aligned', True)
        self.out_channels, self.featmap_strides, self.finest_scale = out_channels, list(featmap_strides), finest_scale
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def map_roi_levels(self, rois, num_levels):
        scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        return torch.floor(torch.log2(scale / self.finest_scale + 1e-6)).clamp(min=0, max=num_levels - 1).long()

    def forward(self, feats, rois):
        from torchvision.ops import roi_align
        feats = feats[:self.num_inputs]
        out = feats[0].new_zeros((rois.size(0), self.out_channels) + self.output_size)
        if rois.size(0) == 0:
            return out
        if len(feats) == 1:
            return roi_align(feats[0], rois, self.output_size, 1.0 / self.featmap_strides[0], self.sampling_ratio, self.aligned)
        lvls = self.map_roi_levels(rois, len(feats))
        for i, stride in enumerate(self.featmap_strides):
            idx = (lvls == i).nonzero(as_tuple=True)[0]
            if idx.numel():
                out[idx] = roi_align(feats[i], rois[idx], self.output_size, 1.0 / stride, self.sampling_ratio, self.aligned)
        return out


def bbox2roi(bbox_list):
    """mmdet.core.bbox2roi: list of (n_i, >=4) boxes per image -> (sum n_i, 5) [batch_idx, x1, y1, x2, y2]."""
    rois = []
    for img_id, bboxes in enumerate(bbox_list):
        ind = bboxes.new_full((bboxes.size(0), 1), img_id)
        rois.append(torch.cat([ind, bboxes[:, :4]], dim=-1))
    return torch.cat(rois, 0) if rois else torch.zeros((0, 5))


def bbox2result(bboxes, labels, num_classes):
    """mmdet.core.bbox2result: (n, 5) boxes + (n,) labels -> list of per-class (k, 5) numpy arrays."""
    import numpy as np
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    bboxes, labels = bboxes.detach().cpu().numpy(), labels.detach().cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]


_OUT_OF_SCOPE = ('bbox_roi_extractor', 'bbox_head', 'global_head', 'noc_roi_extractor',
                 'shared_head', 'mask_roi_extractor', 'mask_head')


@HEADS.register_module()
class MonoRUnRoIHead(nn.Module):
    """ROI head boundary class (monorun_roi_head.py:13-40).  Builds the on-path sub-heads (noc_head,
    projection_head, pose_head); the mmdet-owned pieces that are outside this path's scope (2-D bbox head,
    MC-dropout global extractor, score head, RoI extractors) are kept as their config dicts in ``self.deferred``
    so that the whole ``roi_head`` block of configs/kitti_*.py is accepted.  ``forward_3d`` is the hot sequence
    of ``simple_test`` (:509-534) on tensors the upstream detector stages provide."""

    def __init__(self, noc_roi_extractor=None, noc_head=None, global_head=None, projection_head=None,
                 pose_head=None, score_head=None, debug=False, train_cfg=None, test_cfg=None, **kwargs):
        super(MonoRUnRoIHead, self).__init__()
        self.deferred = {k: v for k, v in dict(kwargs, noc_roi_extractor=noc_roi_extractor,
                                               global_head=global_head).items() if v is not None}
        unknown = set(kwargs) - set(_OUT_OF_SCOPE)
        if unknown:
            raise TypeError(f'unexpected MonoRUnRoIHead arguments: {sorted(unknown)}')
        from .config import ConfigDict
        wrap = lambda c: ConfigDict(c) if isinstance(c, dict) and not isinstance(c, ConfigDict) else c
        self.train_cfg, self.test_cfg, self.debug = wrap(train_cfg), wrap(test_cfg), debug
        self.new_version = True      # mmdet >= 2.4: simple_test returns a list with one dict per image (:601-604)
        self.bbox_stage = None       # the 2-D detection stage of simple_test, see set_bbox_stage
        for name in ('bbox_roi_extractor', 'noc_roi_extractor'):   # caller-side stages simple_test needs
            cfg = self.deferred.get(name)
            if cfg is not None and cfg.get('type') in ROI_EXTRACTORS and len(cfg) > 1:
                setattr(self, name, build_roi_extractor(cfg))
        if global_head is not None and global_head.get('type') in HEADS and len(global_head) > 1:
            self.global_head = build_head(global_head)   # a bare dict(type=...) stub stays deferred
        if noc_head is not None:
            self.noc_head = build_head(noc_head)
        if projection_head is not None:
            self.projection_head = build_head(projection_head)
        if pose_head is not None:
            self.pose_head = build_head(pose_head)
        if score_head is not None:
            self.score_head = build_head(score_head)

    @property
    def with_noc(self):
        return hasattr(self, 'noc_head') and self.noc_head is not None

    @property
    def with_pose(self):
        return hasattr(self, 'pose_head') and self.pose_head is not None

    @property
    def with_score(self):
        return hasattr(self, 'score_head') and self.score_head is not None

    def init_weights(self):
        if hasattr(self, 'global_head'):
            self.global_head.init_weights()
        if self.with_noc:
            self.noc_head.init_weights()
        if self.with_score:
            self.score_head.init_weights()

    def reg_forward(self, reg_feats, det_labels, decode_dims=True):
        """``_reg_forward`` + the decode that follows it (monorun_roi_head.py:489-507) on the 7x7 RoI features:
        returns dict(latent_pred, latent_var, dimensions_pred, dimensions_var, reg_fc_out, dim_pred, dim_var).
        ``decode_dims=False`` leaves ``dimensions_pred`` / ``dimensions_var`` None: the fused PnP entry decodes the
        encoded ``dim_pred`` / ``dim_var`` in its prologue (``forward_3d(..., dim_coder=...)``)."""
        gh = self.global_head
        dim_latent_pred, dim_latent_var, _, _, reg_fc_out = gh(reg_feats)
        dim_pred, dim_var, latent_pred, latent_var = gh.slice_pred(dim_latent_pred, dim_latent_var, det_labels)
        dimensions_pred = dimensions_var = None
        if decode_dims:
            dimensions_pred, dimensions_var = gh.dim_coder.decode(dim_pred, dim_var, det_labels)
        return dict(latent_pred=latent_pred, latent_var=latent_var, dimensions_pred=dimensions_pred,
                    dimensions_var=dimensions_var, reg_fc_out=reg_fc_out, dim_pred=dim_pred, dim_var=dim_var)

    def forward_scores(self, rows, reg_fc_out, dimensions_pred, det_scores=None, cov_correction=True, calib_scoring=False,
                       mult_2d_score=True):
        """monorun_roi_head.py:530-556 on the solver's result rows: covariance correction, score head, sigmoid,
        invalid -> 0, product with the 2-D score.  Returns (scores [N], bbox_3d [N,8], pose_cov_calib [N,4,4])."""
        ph = self.projection_head
        return self.score_head.forward_rows(
            reg_fc_out, rows, dimensions_pred, cov_calib_logscale=self.pose_head.cov_calib_logscale.detach(),
            cov_correction_sd=ph.proj_error_coder.scaling_denomitor if cov_correction else 0.0,
            distance_z_depth=ph.distance_mode == 'z-depth', calib_scoring=calib_scoring,
            det_scores=det_scores if mult_2d_score else None)

    def nms_3d(self, bbox_3d, det_labels, group_offsets=None, nms_thr=None, max_group=None):
        """monorun_roi_head.py:619-655 for all images and classes in one launch: keep mask [N] of the class-wise rotated
        BEV NMS (``nms_thr`` defaults to test_cfg.nms_3d_thr, configs/kitti_multiclass.py:195-210)."""
        from . import pnp
        if nms_thr is None:
            nms_thr = self._test('nms_3d_thr', 0.25)
        return pnp.nms_bev(bbox_3d, det_labels, group_offsets, nms_thr, max_group=max_group)

    def forward_3d(self, noc_feats, bbox_3d_rois, det_labels, latent_pred, dimensions_pred, dimensions_var,
                   cam_intrinsic, img_shape, flip=False, distance_pred=None, cov_correction=True, fused=False,
                   native_head=False, dim_coder=None):
        """monorun_roi_head.py:509-534: dense head -> decode -> analytic coords_2d -> PnP -> covariance correction.

        ``dim_coder`` (fused only): ``dimensions_pred`` / ``dimensions_var`` are the ENCODED regression outputs and
        MultiClassNormDimCoder.decode (:503-507) runs in the PnP prologue; the decoded tensors come back in the dict.

        noc_feats (N,256,14,14), bbox_3d_rois (N,5), det_labels (N,), latent_pred (N,16), dimensions_pred (N,3),
        dimensions_var (N,3)|None, cam_intrinsic (1|N,3,3), img_shape (h, w).
        Returns dict(ret_val, yaw_pred, t_vec_pred, pose_cov_pred, pose_cov_calib, coords_3d, proj_logstd).
        """
        img_shapes = cam_intrinsic.new_tensor(img_shape[:2])[None, ...]
        if fused:  # slice_pred + :513-529 as ONE launch on the head's unsliced output (noc_var is None in this head)
            head = self.noc_head
            all_pred = head.forward_all(noc_feats, latent_pred, flip, native=native_head)
            sliced = head.class_agnostic or head.uncert_channels != 2 or all_pred.stride()[1:] != \
                (all_pred.shape[2] * all_pred.shape[3], all_pred.shape[3], 1)
            if sliced:   # layouts the in-kernel class gather does not cover: slice with torch, still one PnP launch
                noc_pred, _, proj_logstd = head.slice_pred(all_pred, det_labels)
                kw = {}
            else:
                noc_pred, proj_logstd, kw = all_pred, None, dict(labels=det_labels, num_classes=head.num_classes)
            if dim_coder is not None:
                kw.update(dim_coder=dim_coder, dim_labels=det_labels)
            ret_val, yaw, t_vec, cov, cov_calib, *decoded = self.pose_head.forward_fused(
                noc_pred, proj_logstd, bbox_3d_rois, dimensions_pred, dimensions_var, cam_intrinsic, img_shapes,
                head.coord_coder, self.projection_head.proj_error_coder, distance=distance_pred, **kw)
            if cov_correction:
                distance = self.projection_head.get_distance(t_vec)
                cov_calib = self.projection_head.proj_error_coder.cov_correction(cov_calib, distance)
            out = dict(ret_val=ret_val, yaw_pred=yaw, t_vec_pred=t_vec, pose_cov_pred=cov, pose_cov_calib=cov_calib)
            if dim_coder is not None:
                out['dimensions_pred'], out['dimensions_var'] = decoded
            return out
        if native_head:
            noc_pred, noc_var, proj_logstd = self.noc_head.slice_pred(
                self.noc_head.forward_all(noc_feats, latent_pred, flip, native=True), det_labels)
        else:
            noc_pred, noc_var, proj_logstd, _ = self.noc_head(noc_feats, latent_pred, None, det_labels, flip=flip)
        coords_3d, coords_3d_var = self.noc_head.coord_coder.decode(           # :513-515
            noc_pred, noc_var, dimensions_pred, dimensions_var, flip)
        proj_logstd = self.projection_head.proj_error_coder.decode_logstd(     # :516-519
            proj_logstd, coords_3d_var, distance_pred)
        coords_2d_roi = coords_2d_from_rois(bbox_3d_rois, noc_pred.shape[-1])  # :521-523
        ret_val, yaw, t_vec, cov, cov_calib = self.pose_head(                  # :525-529
            coords_2d_roi, proj_logstd, coords_3d, cam_intrinsic, img_shapes)
        if cov_correction:                                                     # :530-534
            distance = self.projection_head.get_distance(t_vec)
            cov_calib = self.projection_head.proj_error_coder.cov_correction(cov_calib, distance)
        return dict(ret_val=ret_val, yaw_pred=yaw, t_vec_pred=t_vec, pose_cov_pred=cov, pose_cov_calib=cov_calib,
                    coords_3d=coords_3d, proj_logstd=proj_logstd, coords_2d=coords_2d_roi)

    # ------------------------------------------------------------------ simple_test (monorun_roi_head.py:442-605)
    def _test(self, key, default=None):
        cfg = self.test_cfg
        if cfg is None:
            return default
        return cfg.get(key, default) if isinstance(cfg, dict) else getattr(cfg, key, default)

    def set_bbox_stage(self, fn):
        """The 2-D detection stage of ``simple_test`` (:460-469: ``_bbox_forward`` + ``bbox_head.get_bboxes``) belongs
        to mmdet (Shared2FCBBoxHead, its coder and NMS) and is outside this path; it is injected instead:
        ``fn(x, proposal_list, img_metas, rescale, test_cfg) -> (det_bboxes (n, 5) [x1, y1, x2, y2, score],
        det_labels (n,) int64)`` -- an mmdet ``StandardRoIHead.simple_test_bboxes`` bound method has this shape."""
        self.bbox_stage = fn
        return self

    @property
    def with_bbox(self):
        return self.bbox_stage is not None

    def _reg_forward(self, x, rois, labels, decode_dims=True):
        """:272-288 plus the decode that follows it in simple_test (:503-507)."""
        reg_feats = self.bbox_roi_extractor(x[:self.bbox_roi_extractor.num_inputs], rois)
        out = self.reg_forward(reg_feats, labels, decode_dims=decode_dims)
        out['roi_feats'] = reg_feats
        return out

    def get_bbox_3d_result(self, dimensions, yaw, t_vec, scores, labels, to_np=False):
        """:607-613: per-class lists of [l, h, w, x, y, z, ry, score] rows."""
        bboxes_3d = torch.cat((dimensions, t_vec, yaw, scores.unsqueeze(1)), dim=1)
        if to_np:
            bboxes_3d, labels = bboxes_3d.cpu().numpy(), labels.cpu().numpy()
        return [bboxes_3d[labels == i] for i in range(self.noc_head.num_classes)]

    def multiclass_3d_result_nms(self, bbox_3d_result, nms_thr=0.25, to_np=True):
        """:619-655 -- per class, rotated-BEV NMS; returns (kept boxes, kept indices into the class's rows), both in
        descending score order like mmdet3d's ``nms_gpu``.  One ``mrpnp_nms_bev`` launch covers all classes."""
        from . import pnp
        sizes = [int(b.size(0)) for b in bbox_3d_result]
        out_boxes, out_inds = [], []
        keep_all = None
        if sum(sizes) > 0:
            allb = torch.cat(bbox_3d_result, 0)
            labels = torch.cat([allb.new_full((k,), i, dtype=torch.long) for i, k in enumerate(sizes)])
            keep_all = pnp.nms_bev(allb, labels, None, nms_thr)
        start = 0
        for boxes, k in zip(bbox_3d_result, sizes):
            if k > 1:
                keep = keep_all[start:start + k]
                inds = keep.nonzero(as_tuple=True)[0]
                inds = inds[torch.argsort(boxes[inds, 7], descending=True, stable=True)]
                kept = boxes[inds]
            else:   # :645-654: zero or one box is returned as is, with indices zeros(n)
                inds, kept = boxes.new_zeros((k,), dtype=torch.int64), boxes
            out_boxes.append(kept.cpu().numpy() if to_np else kept)
            out_inds.append(inds.cpu().numpy() if to_np else inds)
            start += k
        return out_boxes, out_inds

    def simple_test(self, x, proposal_list, img_metas, proposals=None, coord_2d=None, cam_intrinsic=None, rescale=False,
                    native=True):
        """Drop-in for ``MonoRUnRoIHead.simple_test`` (monorun_roi_head.py:442-605): one image in, ``[dict(bbox_results,
        bbox_3d_results)]`` out (per-class numpy arrays; 3-D rows [l, h, w, x, y, z, ry, score]).

        The 2-D stage is the injected ``bbox_stage``.  From its detections on, everything is this repo's native
        sequence: RoI features -> MC-dropout global extractor -> dense head (tcgen05 convolutions) -> fused decode +
        uncertainty PnP (one launch, incl. the reprojection-threshold consensus of ``epnp_ransac_thres_ratio``) -> score
        stage -> 3-D NMS.  ``coord_2d`` is accepted for signature compatibility: the RoI pixel grid is generated
        analytically from the boxes, which assumes the untransformed pixel grid of the shipped test
        pipelines (scale_factor 1, no flip); other metas fall back to resampling ``coord_2d`` with ``roi_align``.
        ``native=False`` runs the fp32 torch modules instead of the kernels (the reference of the tests)."""
        import numpy as np
        assert self.with_bbox and self.with_noc and self.with_pose and self.with_score, \
            'simple_test needs the bbox stage (set_bbox_stage) and the noc / pose / score heads'
        assert len(img_metas) == 1, 'batch inference is not supported yet'   # :452
        meta = img_metas[0]
        img_shape, scale_factor = meta['img_shape'], meta.get('scale_factor', 1.0)
        flip = bool(meta.get('flip', False))
        cam_intrinsic = cam_intrinsic[0][0][None, ...]                        # :458
        num_classes = self.noc_head.num_classes

        det_bboxes, det_labels = self.bbox_stage(x, proposal_list, img_metas, rescale, self.test_cfg)   # :460-469
        if det_bboxes.shape[0] > 0:                                           # :471-478
            sf = det_bboxes.new_tensor(scale_factor) if not isinstance(scale_factor, float) else scale_factor
            _bboxes = det_bboxes[:, :4] * sf if rescale else det_bboxes
            bbox_3d_rois = bbox2roi([_bboxes])
        else:
            bbox_3d_rois = None
        bbox_result = bbox2result(det_bboxes, det_labels, num_classes)       # :481-482
        if bbox_3d_rois is None:                                              # :485-487
            bbox_3d_result = [np.zeros((0, 8), dtype=np.float32) for _ in range(num_classes)]
            return [dict(bbox_results=bbox_result, bbox_3d_results=bbox_3d_result)]

        with torch.no_grad():
            unit_scale = isinstance(scale_factor, float) and scale_factor == 1.0 or \
                (not isinstance(scale_factor, float) and bool((np.asarray(scale_factor) == 1).all()))
            fused = native and x[0].is_cuda and not flip and unit_scale
            # fused: the dimension decode of :503-507 runs in the PnP prologue, on the encoded regression output
            reg = self._reg_forward(x, bbox_3d_rois, det_labels, decode_dims=not fused)   # :489-507
            noc_feats = self.noc_roi_extractor(x[:self.noc_roi_extractor.num_inputs], bbox_3d_rois)   # :331-332
            if fused:
                out = self.forward_3d(noc_feats, bbox_3d_rois, det_labels, reg['latent_pred'], reg['dim_pred'],
                                      reg['dim_var'], cam_intrinsic, img_shape, flip=flip, cov_correction=False,
                                      fused=True, native_head=True, dim_coder=self.global_head.dim_coder)
                reg['dimensions_pred'], reg['dimensions_var'] = out['dimensions_pred'], out['dimensions_var']
            elif coord_2d is None:
                out = self.forward_3d(noc_feats, bbox_3d_rois, det_labels, reg['latent_pred'], reg['dimensions_pred'],
                                      reg['dimensions_var'], cam_intrinsic, img_shape, flip=flip,
                                      cov_correction=False, fused=False, native_head=False)
            else:   # transformed pixel grid: resample it like the reference does (:521-523)
                out = self._forward_3d_resampled(noc_feats, bbox_3d_rois, det_labels, reg, cam_intrinsic, img_shape, flip,
                                                 coord_2d[0])
            n = det_labels.numel()
            rows = torch.cat([out['yaw_pred'], out['t_vec_pred'], out['pose_cov_pred'].reshape(n, 16),
                              out['ret_val'].float()[:, None], out['yaw_pred'].new_zeros(n, 3)], 1)
            scores, bbox_3d, _ = self.forward_scores(                         # :530-550
                rows, reg['reg_fc_out'], reg['dimensions_pred'], det_scores=det_bboxes[:, -1],
                cov_correction=bool(self._test('cov_correction', False)), calib_scoring=bool(self._test('calib_scoring', False)),
                mult_2d_score=bool(self._test('mult_2d_score', False)))
            per_class = [bbox_3d[det_labels == i] for i in range(num_classes)]            # :552-558
            bbox_3d_result, keep_inds_3d = self.multiclass_3d_result_nms(per_class, self._test('nms_3d_thr', 0.25), to_np=True)
        bbox_result = [b[k] for b, k in zip(bbox_result, keep_inds_3d)]      # :561-565
        return [dict(bbox_results=bbox_result, bbox_3d_results=bbox_3d_result)]

    def _forward_3d_resampled(self, noc_feats, rois, det_labels, reg, cam_intrinsic, img_shape, flip, coord_2d):
        """The unfused sequence with ``coords_2d_roi = roi_align(coord_2d, rois, ...)`` (:513-529) for pixel grids that went
        through Resize / Flip / Pad."""
        from torchvision.ops import roi_align
        noc_pred, noc_var, proj_logstd, _ = self.noc_head(noc_feats, reg['latent_pred'], None, det_labels, flip=flip)
        coords_3d, coords_3d_var = self.noc_head.coord_coder.decode(noc_pred, noc_var, reg['dimensions_pred'],
                                                                    reg['dimensions_var'], flip)
        proj_logstd = self.projection_head.proj_error_coder.decode_logstd(proj_logstd, coords_3d_var, None)
        coords_2d_roi = roi_align(coord_2d[None] if coord_2d.dim() == 3 else coord_2d, rois, noc_pred.shape[-2:], 1.0, 0, True)
        img_shapes = cam_intrinsic.new_tensor(img_shape[:2])[None, ...]
        ret_val, yaw, t_vec, cov, cov_calib = self.pose_head(coords_2d_roi, proj_logstd, coords_3d, cam_intrinsic, img_shapes)
        return dict(ret_val=ret_val, yaw_pred=yaw, t_vec_pred=t_vec, pose_cov_pred=cov, pose_cov_calib=cov_calib)
